#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/microbench_cg.py iso2_512 10 2 > gpurun_out/r2_cg_iso2_nd.log 2>&1; tail -3 gpurun_out/r2_cg_iso2_nd.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_iso2_nd.csv python scripts/microbench_cg.py iso2_512 3 1 > gpurun_out/r2_ncu_iso2_nd.log 2>&1; tail -1 gpurun_out/r2_ncu_iso2_nd.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize_oracle.py > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_all.log
tail -12 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench.log
tail -3 gpurun_out/r2_bench.log
