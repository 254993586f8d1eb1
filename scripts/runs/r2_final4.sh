#!/bin/bash
# last validation of the round: every GPU test, smoke, rotated bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu6.log
tail -3 gpurun_out/r2_pytest_gpu6.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke4.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke4.log
tail -2 gpurun_out/r2_smoke4.log
timeout 600 python bench.py --workload sr3_256_rigid --no-sharded --steps 3 > gpurun_out/r2_bench_sr3_256_rigid3.log 2>&1; grep '^{' gpurun_out/r2_bench_sr3_256_rigid3.log | tail -1 | cut -c1-200
