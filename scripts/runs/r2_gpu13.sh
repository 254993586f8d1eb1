#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "several_observations or rotated or multi_axis" > gpurun_out/r2_pytest_multiobs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multiobs.log
tail -15 gpurun_out/r2_pytest_multiobs.log
