#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_midsize.py -m gpu -q -x > gpurun_out/r2_pytest_exfuse.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_exfuse.log
tail -3 gpurun_out/r2_pytest_exfuse.log
for f in 0 1; do
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 rot_expand_fused=$f > gpurun_out/r2_cg_rigid_exfuse$f.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_exfuse$f.log | cut -c1-60
done
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rot_forward|rot_adjoint_cell" -s 20 -c 2 --csv --log-file gpurun_out/r2_exfuse_inst.csv python scripts/microbench_cg.py sr3_256_rigid 20 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_exfuse_inst.csv | awk -F'","' '{print substr($5,1,40), $13, $15}' | cut -c1-150
