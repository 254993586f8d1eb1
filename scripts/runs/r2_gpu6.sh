#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/microbench_cg.py iso2_512 10 2 > gpurun_out/r2_cg_iso2_nd.log 2>&1; tail -3 gpurun_out/r2_cg_iso2_nd.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_iso2_nd.csv python scripts/microbench_cg.py iso2_512 3 1 > gpurun_out/r2_ncu_iso2_nd.log 2>&1; tail -1 gpurun_out/r2_ncu_iso2_nd.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_oracle.py > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_all.log
tail -12 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench.log
tail -3 gpurun_out/r2_bench.log
timeout 600 python scripts/r2_sweep_fast.py sr3_256 > gpurun_out/r2_sweep_fast.log 2>&1; sort -t: -k2 -n gpurun_out/r2_sweep_fast.log | head -5
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jtv_kernel_vec|rhs_fused_kernel|nd_down_kernel|nd_up_kernel" -c 8 -o gpurun_out/r2_prof_admm python scripts/microbench_admm.py sr3_256 > gpurun_out/r2_ncu_admm.log 2>&1; tail -2 gpurun_out/r2_ncu_admm.log
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nd_down_kernel|nd_up_kernel" -s 2 -c 2 -o gpurun_out/r2_prof_nd python scripts/microbench_cg.py iso2_512 2 1 > gpurun_out/r2_ncu_nd.log 2>&1; tail -2 gpurun_out/r2_ncu_nd.log
