#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_fit.py -m gpu -q 2>&1 | tail -5
(for wl in sr3_256 thickz2_256; do timeout 300 python scripts/microbench_admm.py $wl 2>&1 | tail -5; done) 2>&1 | tee gpurun_out/admm_pieces3.log
