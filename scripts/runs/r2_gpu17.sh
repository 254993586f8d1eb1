#!/bin/bash
# cell-coefficient adjoint of rotated operators: parity, then timing against the gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "rotated or adjoint or multi" > gpurun_out/r2_pytest_cell.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell.log
tail -15 gpurun_out/r2_pytest_cell.log
for cell in 0 1; do
  NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 rot_cell=$cell > gpurun_out/r2_cg_rigid_cell$cell.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_cell$cell.log
done
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256 20 5 > gpurun_out/r2_cg_sr3_to7.log 2>&1; tail -3 gpurun_out/r2_cg_sr3_to7.log
