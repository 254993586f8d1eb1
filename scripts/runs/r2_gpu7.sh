#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stream_kernel.py tests/test_gpu_solver.py -m gpu -q -x -k "multi_axis or iso2 or golden or graph or jtv or admm" > gpurun_out/r2_pytest_nd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_nd.log
tail -15 gpurun_out/r2_pytest_nd.log
timeout 600 python -m pytest tests/test_gpu_fullsize_oracle.py -m gpu -q -k "iso2" > gpurun_out/r2_pytest_nd2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_nd2.log
tail -5 gpurun_out/r2_pytest_nd2.log
timeout 300 python scripts/microbench_cg.py iso2_512 10 2 > gpurun_out/r2_cg_iso2_nd.log 2>&1; tail -3 gpurun_out/r2_cg_iso2_nd.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_iso2_nd.csv python scripts/microbench_cg.py iso2_512 3 1 > gpurun_out/r2_ncu_iso2_nd.log 2>&1; tail -1 gpurun_out/r2_ncu_iso2_nd.log
timeout 300 python scripts/microbench_admm.py iso2_512 > gpurun_out/r2_admm_iso2.log 2>&1; tail -8 gpurun_out/r2_admm_iso2.log
timeout 300 python scripts/microbench_admm.py sr3_256 > gpurun_out/r2_admm_sr3.log 2>&1; tail -8 gpurun_out/r2_admm_sr3.log
timeout 300 python scripts/microbench_admm.py sr3_256 jtv_wide=1 > gpurun_out/r2_admm_sr3_wide.log 2>&1; tail -3 gpurun_out/r2_admm_sr3_wide.log
