#!/bin/bash
# cell-coefficient adjoint of rotated operators: parity, then timing against the gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "rotated or adjoint or multi" > gpurun_out/r2_pytest_cell2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell2.log
tail -15 gpurun_out/r2_pytest_cell2.log
for cell in 1; do
  NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 rot_cell=$cell > gpurun_out/r2_cg_rigid_cellv2_$cell.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_cellv2_$cell.log
done
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rot_adjoint_cell|rot_forward" -s 10 -c 4 --csv --log-file gpurun_out/r2_launches_rigid_cell2.csv python scripts/microbench_cg.py sr3_256_rigid 20 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_launches_rigid_cell2.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-150
