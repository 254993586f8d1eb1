#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
(echo "== iso2_512"; timeout 600 python scripts/microbench_admm.py iso2_512 2>&1 | tail -5; timeout 600 python scripts/microbench_cg.py iso2_512 20 2 2>&1 | tail -3 | cut -c1-110) 2>&1 | tee gpurun_out/admm_iso2b.log
