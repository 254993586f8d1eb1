#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"nd_up_spec_kernel<0" -s 1 -c 1 -o gpurun_out/r2_prof_ndup_plain python scripts/microbench_cg.py iso2_512 2 1 > gpurun_out/r2_ncu_ndup.log 2>&1; tail -1 gpurun_out/r2_ncu_ndup.log
