#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) full captures of the fused
# (COMBINE) lean matvec of each channel of sr3_256 (thick-x, thick-y, thick-z), (3) residual update
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -1 gpurun_out/bench_ncu.log | cut -c1-100
for c in 0 1 2; do
  s=$((8 + 43 * c))
  ncu --set full --clock-control none --import-source on -k regex:lhs_fast_kernel -s $s -c 1 -o gpurun_out/prof_fast_ch$c python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/ncu_run$c.log 2>&1
  tail -1 gpurun_out/ncu_run$c.log | cut -c1-100
done
