#!/bin/bash
# ncu: full capture of the fused (COMBINE) lean matvec, thick-x channel (8th lhs launch of the
# first solve) and thick-z channel (channel 2)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lhs_fast_kernel -s 8 -c 1 -o gpurun_out/prof_fast_m python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
ncu --set full --clock-control none --import-source on -k regex:lhs_fast_kernel -s 96 -c 1 -o gpurun_out/prof_fast_z python scripts/microbench_cg.py sr3_256 20 1 > gpurun_out/ncu_run2.log 2>&1
tail -3 gpurun_out/ncu_run2.log
