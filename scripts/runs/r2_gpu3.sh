#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rot_adjoint_kernel|rot_forward_kernel|lhs_direct_kernel" -s 2 -c 6 -o gpurun_out/r2_prof_rot python scripts/microbench_cg.py sr3_256_rigid 3 1 > gpurun_out/r2_ncu_rot.log 2>&1; tail -3 gpurun_out/r2_ncu_rot.log
