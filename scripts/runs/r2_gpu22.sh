#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "rotated or adjoint or multi" > gpurun_out/r2_pytest_cell4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell4.log
tail -4 gpurun_out/r2_pytest_cell4.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_cellv4.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_cellv4.log
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:"rot_adjoint_cell" -s 10 -c 1 --csv --log-file gpurun_out/r2_launches_rigid_cell4.csv python scripts/microbench_cg.py sr3_256_rigid 20 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_launches_rigid_cell4.csv | awk -F'","' '{print $13, $15}' | cut -c1-150
