#!/bin/bash
# round 2, GPU pass 2: rotated fused kernels -- parity, then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_midsize.py tests/test_compat_dropin.py -m gpu -q -x > gpurun_out/r2_pytest_rot.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_rot.log
tail -15 gpurun_out/r2_pytest_rot.log
timeout 600 python -m pytest tests/test_gpu_fullsize_oracle.py tests/test_gpu_solver.py -m gpu -q -k "rigid" > gpurun_out/r2_pytest_rot2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_rot2.log
tail -15 gpurun_out/r2_pytest_rot2.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_fused.log 2>&1; tail -4 gpurun_out/r2_cg_rigid_fused.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_rigid_fused.csv python scripts/microbench_cg.py sr3_256_rigid 5 1 > gpurun_out/r2_ncu_rigid_fused.log 2>&1; tail -2 gpurun_out/r2_ncu_rigid_fused.log
