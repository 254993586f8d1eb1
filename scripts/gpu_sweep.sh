#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_stream_kernel.py -m gpu -q -k fuzz 2>&1 | tail -30
