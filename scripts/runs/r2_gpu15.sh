#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rhs_fused4_kernel|nll_nd_kernel" -c 4 -o gpurun_out/r2_prof_rhs python scripts/microbench_admm.py sr3_256 > gpurun_out/r2_ncu_rhs.log 2>&1; tail -1 gpurun_out/r2_ncu_rhs.log
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jtv_kernel_vec" -s 6 -c 2 -o gpurun_out/r2_prof_jtv python scripts/microbench_admm.py sr3_256 > gpurun_out/r2_ncu_jtv.log 2>&1; tail -1 gpurun_out/r2_ncu_jtv.log
