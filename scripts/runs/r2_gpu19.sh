#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rot_adjoint_cell_kernel" -s 5 -c 1 -o gpurun_out/r2_prof_rotcell python scripts/microbench_cg.py sr3_256_rigid 20 1 > gpurun_out/r2_ncu_rotcell.log 2>&1; tail -2 gpurun_out/r2_ncu_rotcell.log
