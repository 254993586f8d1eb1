#!/bin/bash
mkdir -p gpurun_out
(for wl in thick3_256 thick5_256 thick6_256; do for v in 0 2; do echo "== $wl lhs_variant $v"; timeout 200 python scripts/microbench_cg.py $wl 20 3 lhs_variant=$v 2>&1 | tail -3 | cut -c1-100; done; done) | tee gpurun_out/ratios.log
