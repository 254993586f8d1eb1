#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python scripts/microbench_multiview.py 2>&1 | tail -6 | tee gpurun_out/multiview.log
