#!/bin/bash
# ncu: full capture of the streaming kernel (thick-x channel) + launch list of one bench step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lhs_stream -s 2 -c 1 -o gpurun_out/prof_stream python scripts/microbench_lhs.py sr3_256 0 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log | cut -c1-300
