#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
(for f in 1 0; do echo "== cg_fuse $f"; timeout 300 python scripts/microbench_admm.py sr3_256 1e-3 cg_fuse=$f 2>&1 | grep _update_admm; done) 2>&1 | tee gpurun_out/efuse.log
