#!/bin/bash
# final numbers with the final defaults (L2 hints off): tests, bench lines, launch list, lattice configs
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu4.log
tail -3 gpurun_out/r2_pytest_gpu4.log
timeout 900 python bench.py > gpurun_out/r2_bench_final3.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_final3.log
tail -2 gpurun_out/r2_bench_final3.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench3.csv python bench.py --steps 2 --warmup 1 --no-sharded --no-cpu-baseline > gpurun_out/r2_ncu_bench3.log 2>&1; echo "ncu rc=$?"
out=gpurun_out/r2_configs_lattice.txt; : > $out
for w in denoise_181 sr3_256 crop3_256 thickz2_256 thickz2_384; do
  echo "== $w" >> $out
  timeout 300 python scripts/microbench_cg.py $w 20 3 >> $out 2>&1
done
for w in sr3_256 denoise_181; do
  echo "== admm pieces $w" >> $out
  timeout 300 python scripts/microbench_admm.py $w >> $out 2>&1
done
cat $out | cut -c1-170
timeout 600 python bench.py --workload sr3_256_rigid --no-sharded --steps 3 > gpurun_out/r2_bench_sr3_256_rigid2.log 2>&1; tail -1 gpurun_out/r2_bench_sr3_256_rigid2.log | cut -c1-300
