#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_midsize.py -m gpu -q -x > gpurun_out/r2_pytest_fwd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_fwd.log
tail -3 gpurun_out/r2_pytest_fwd.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_fwd.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_fwd.log | cut -c1-60
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rot_forward" -s 10 -c 1 --csv --log-file gpurun_out/r2_fwd_inst.csv python scripts/microbench_cg.py sr3_256_rigid 20 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_fwd_inst.csv | awk -F'","' '{print $13, $15}' | cut -c1-150
