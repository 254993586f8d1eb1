#!/bin/bash
# final validation of round 2 (second session): every GPU test, smoke, bench lines, launch list,
# full capture of the fused matvec (traffic), per-config numbers
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu3.log
tail -12 gpurun_out/r2_pytest_gpu3.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke3.log
tail -2 gpurun_out/r2_smoke3.log
timeout 900 python bench.py > gpurun_out/r2_bench_final2.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_final2.log
tail -2 gpurun_out/r2_bench_final2.log | cut -c1-700
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference2.log 2>&1; echo "ref rc=$?" >> gpurun_out/r2_bench_reference2.log
tail -2 gpurun_out/r2_bench_reference2.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench2.csv python bench.py --steps 2 --warmup 1 --no-sharded > gpurun_out/r2_ncu_bench2.log 2>&1; echo "ncu rc=$?"
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lhs_fast_kernel" -s 12 -c 3 -o gpurun_out/r2_prof_fast_final python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/r2_ncu_fast_final.log 2>&1; tail -1 gpurun_out/r2_ncu_fast_final.log
bash scripts/r2_configs.sh > gpurun_out/r2_configs_run2.log 2>&1; tail -5 gpurun_out/r2_configs_run2.log | cut -c1-300
