#!/bin/bash
# balanced z tiles of the lean kernel (run-time tz): parity, denoise_181 / sr3_256 / thickz2_384 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_kernel.py tests/test_gpu_midsize.py -m gpu -q -x > gpurun_out/r2_pytest_tz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_tz.log
tail -3 gpurun_out/r2_pytest_tz.log
for t in 128 0; do
timeout 300 python scripts/microbench_cg.py denoise_181 20 5 fast_tz=$t 2>&1 | tail -1 | cut -c1-110
done
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256 20 3 2>&1 | tail -3 | cut -c1-60
