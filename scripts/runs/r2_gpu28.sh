#!/bin/bash
# full captures of the fused (COMBINE) lean matvec of each channel of sr3_256 with the final defaults
mkdir -p gpurun_out
for c in 0 1 2; do
  s=$((8 + 43 * c))
  NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lhs_fast_kernel -s $s -c 1 -o gpurun_out/r2_prof_fast_final_ch$c python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/r2_ncu_fast_final_ch$c.log 2>&1
  tail -1 gpurun_out/r2_ncu_fast_final_ch$c.log | cut -c1-100
done
