#!/bin/bash
mkdir -p gpurun_out
(for r in 8 4 2; do echo "== jtv_rows $r"; timeout 300 python scripts/microbench_admm.py sr3_256 1e-3 jtv_rows=$r 2>&1 | grep "jtv prox"; done) | tee gpurun_out/jtv_rows.log
