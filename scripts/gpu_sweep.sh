#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
(for wl in denoise_181; do timeout 300 python scripts/microbench_admm.py $wl 2>&1 | tail -5; done
echo "== cg denoise_181"; timeout 200 python scripts/microbench_cg.py denoise_181 20 5 2>&1 | tail -1
) 2>&1 | tee gpurun_out/admm_pieces2.log
