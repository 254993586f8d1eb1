#!/bin/bash
mkdir -p gpurun_out
(for wl in denoise_181 thickz2_256 thickz2_384; do echo "== $wl (throughput mode, 20 fixed CG iterations)"; NOPROF=1 timeout 300 python scripts/microbench_cg.py $wl 20 3 2>&1 | tail -3 | cut -c1-60; timeout 300 python scripts/microbench_cg.py $wl 20 3 2>&1 | tail -3 | cut -c60-130; done
echo "== iso2_512 (general path), one reference-settings ADMM iteration"; timeout 600 python scripts/microbench_admm.py iso2_512 2>&1 | tail -4
) 2>&1 | tee gpurun_out/configs.log
