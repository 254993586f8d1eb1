#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_midsize.py tests/test_gpu_fit.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2_pytest_rhs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_rhs.log
tail -4 gpurun_out/r2_pytest_rhs.log
timeout 300 python scripts/microbench_admm.py sr3_256 > gpurun_out/r2_admm_sr3.log 2>&1; tail -5 gpurun_out/r2_admm_sr3.log
timeout 300 python scripts/microbench_admm.py thickz2_256 > gpurun_out/r2_admm_tz2.log 2>&1; tail -5 gpurun_out/r2_admm_tz2.log
