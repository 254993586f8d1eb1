#!/bin/bash
# round 2, GPU pass 1: new parity tests (full size, mid size, reference-over-compat), then the
# rotated-operator workload and the 512^3 leads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
nproc >> gpurun_out/r2_gpu.txt; free -g | head -2 >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_midsize.py tests/test_compat_dropin.py tests/test_gpu_fullsize_oracle.py -m gpu -q --durations=15 > gpurun_out/r2_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_new.log
tail -30 gpurun_out/r2_pytest_new.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid.log 2>&1; tail -4 gpurun_out/r2_cg_rigid.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_rigid.csv python scripts/microbench_cg.py sr3_256_rigid 5 1 > gpurun_out/r2_ncu_rigid.log 2>&1; tail -2 gpurun_out/r2_ncu_rigid.log
for k in fast_rpt=0 fast_rpt=2; do
  timeout 300 python scripts/microbench_cg.py iso2_512 10 2 $k > gpurun_out/r2_cg_iso2_$k.log 2>&1; echo $k; tail -1 gpurun_out/r2_cg_iso2_$k.log
done
