#!/bin/bash
mkdir -p gpurun_out
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r2_launches_rigid_cell.csv python scripts/microbench_cg.py sr3_256_rigid 20 1 > gpurun_out/r2_ncu_rigid_cell.log 2>&1; tail -2 gpurun_out/r2_ncu_rigid_cell.log
