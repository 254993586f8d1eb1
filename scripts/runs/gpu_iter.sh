#!/bin/bash
# quick iteration: lhs kernel tests, then the CG microbenchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_kernel.py -m gpu -x -q > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iter.log
tail -15 gpurun_out/pytest_iter.log
timeout 300 python scripts/microbench_cg.py sr3_256 20 5 "$@" 2>&1 | tee gpurun_out/microbench_cg.log
