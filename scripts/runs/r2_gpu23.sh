#!/bin/bash
# full validation after the cell adjoint / 7-row tiles: every GPU test, smoke, default bench, admm pieces
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu2.log
tail -14 gpurun_out/r2_pytest_gpu2.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke2.log
tail -2 gpurun_out/r2_smoke2.log
timeout 900 python bench.py > gpurun_out/r2_bench_v2.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_v2.log
tail -2 gpurun_out/r2_bench_v2.log | cut -c1-1500
NOPROF=1 timeout 300 python scripts/microbench_admm.py sr3_256_rigid > gpurun_out/r2_admm_rigid_v2.log 2>&1; tail -12 gpurun_out/r2_admm_rigid_v2.log
