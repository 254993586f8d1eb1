#!/bin/bash
# quick iteration: stream-kernel tests, then microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream_kernel.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iter.log
tail -12 gpurun_out/pytest_iter.log
timeout 300 python scripts/microbench_lhs.py sr3_256 "$@" 2>&1 | tee gpurun_out/microbench.log
