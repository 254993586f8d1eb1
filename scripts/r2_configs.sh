#!/bin/bash
# per-config CG timing (single stream, 20 fixed iterations) + ADMM pieces, for DESIGN.md section 6
mkdir -p gpurun_out
out=gpurun_out/r2_configs.txt; : > $out
for w in denoise_181 sr3_256 crop3_256 thickz2_256 thickz2_384 sr3_256_rigid iso2_512; do
  echo "== $w" >> $out
  timeout 300 python scripts/microbench_cg.py $w 20 3 >> $out 2>&1
done
echo "== denoise_181 cg_graph=0" >> $out
timeout 300 python scripts/microbench_cg.py denoise_181 20 3 cg_graph=0 >> $out 2>&1
for w in sr3_256 sr3_256_rigid iso2_512 denoise_181; do
  echo "== admm pieces $w" >> $out
  timeout 300 python scripts/microbench_admm.py $w >> $out 2>&1
done
cat $out | grep -v "^$" | cut -c1-200
for w in denoise_181 sr3_256_rigid iso2_512; do
  timeout 600 python bench.py --workload $w --no-sharded --steps 3 > gpurun_out/r2_bench_$w.log 2>&1; tail -1 gpurun_out/r2_bench_$w.log | cut -c1-400
done
