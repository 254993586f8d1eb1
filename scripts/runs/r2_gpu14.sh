#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_midsize.py tests/test_gpu_fit.py tests/test_compat_dropin.py -m gpu -q -x > gpurun_out/r2_pytest_nll.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_nll.log
tail -4 gpurun_out/r2_pytest_nll.log
for w in sr3_256 thickz2_256; do timeout 300 python scripts/microbench_admm.py $w > gpurun_out/r2_admm_$w.log 2>&1; tail -5 gpurun_out/r2_admm_$w.log; done
