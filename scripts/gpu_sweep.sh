#!/bin/bash
mkdir -p gpurun_out
(echo "== profile events on"; timeout 120 python scripts/microbench_cg.py sr3_256 20 5 2>&1 | tail -3
echo "== profile events off"; NOPROF=1 timeout 120 python scripts/microbench_cg.py sr3_256 20 5 2>&1 | tail -3
echo "== depth 2"; NOPROF=1 timeout 120 python scripts/microbench_cg.py sr3_256 20 5 fast_depth=2 2>&1 | tail -3
) | tee gpurun_out/sweep3.log
