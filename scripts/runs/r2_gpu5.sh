#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_midsize.py tests/test_gpu_solver.py tests/test_gpu_stream_kernel.py -m gpu -q -x > gpurun_out/r2_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_a.log
tail -25 gpurun_out/r2_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_fullsize_oracle.py -m gpu -q -k "rigid or iso2" > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_b.log
tail -8 gpurun_out/r2_pytest_b.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_fused.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_fused.log
timeout 300 python scripts/microbench_cg.py iso2_512 10 2 > gpurun_out/r2_cg_iso2_nd.log 2>&1; tail -3 gpurun_out/r2_cg_iso2_nd.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_iso2_nd.csv python scripts/microbench_cg.py iso2_512 5 1 > gpurun_out/r2_ncu_iso2_nd.log 2>&1; tail -2 gpurun_out/r2_ncu_iso2_nd.log
