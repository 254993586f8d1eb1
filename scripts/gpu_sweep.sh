#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
(timeout 300 python scripts/microbench_admm.py sr3_256 2>&1 | tail -5) 2>&1 | tee gpurun_out/admm_pieces4.log
