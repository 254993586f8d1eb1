#!/bin/bash
# ncu: full capture of the fused (COMBINE) streaming matvec, thick-x channel and thick-z channel
mkdir -p gpurun_out
python scripts/microbench_cg.py sr3_256 2>&1 | tee gpurun_out/microbench_cg.log
ncu --set full --clock-control none --import-source on -k regex:lhs_stream_kernel -s 8 -c 1 -o gpurun_out/prof_combine_m python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
ncu --set full --clock-control none --import-source on -k regex:lhs_stream_kernel -s 140 -c 1 -o gpurun_out/prof_combine_z python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/ncu_run2.log 2>&1
tail -3 gpurun_out/ncu_run2.log
