#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_solver.py -m gpu -q 2>&1 | tail -25
