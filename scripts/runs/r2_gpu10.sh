#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rot_forward_kernel|lhs_rot_kernel" -s 4 -c 2 -o gpurun_out/r2_prof_rot2 python scripts/microbench_cg.py sr3_256_rigid 3 1 > gpurun_out/r2_ncu_rot2.log 2>&1; tail -1 gpurun_out/r2_ncu_rot2.log
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nd_up_spec_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_ndup_plain python scripts/microbench_cg.py iso2_512 2 1 > gpurun_out/r2_ncu_ndup.log 2>&1; tail -1 gpurun_out/r2_ncu_ndup.log
