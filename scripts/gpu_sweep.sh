#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
(echo "== iso2_512"; timeout 600 python scripts/microbench_admm.py iso2_512 2>&1 | tail -5) 2>&1 | tee gpurun_out/admm_iso2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_iso2.csv python scripts/lhs_once.py iso2_512 > gpurun_out/lhs_once.log 2>&1
tail -1 gpurun_out/lhs_once.log
