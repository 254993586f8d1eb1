#!/bin/bash
# full validation: every GPU test, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -14 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_final.log
tail -2 gpurun_out/r2_bench_final.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference.log 2>&1; echo "ref rc=$?" >> gpurun_out/r2_bench_reference.log
tail -2 gpurun_out/r2_bench_reference.log | cut -c1-400
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_final.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_final.log
