#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stream_kernel.py -m gpu -q -x -k "multi_axis or iso2" > gpurun_out/r2_pytest_nd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_nd.log
tail -3 gpurun_out/r2_pytest_nd.log
timeout 300 python scripts/microbench_cg.py iso2_512 10 2 > gpurun_out/r2_cg_iso2_nd.log 2>&1; tail -3 gpurun_out/r2_cg_iso2_nd.log
NOPROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_iso2_nd.csv python scripts/microbench_cg.py iso2_512 3 1 > gpurun_out/r2_ncu_iso2_nd.log 2>&1; tail -1 gpurun_out/r2_ncu_iso2_nd.log
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nd_down_spec_kernel|nd_up_spec_kernel" -s 3 -c 2 -o gpurun_out/r2_prof_nd_spec python scripts/microbench_cg.py iso2_512 2 1 > gpurun_out/r2_ncu_nd_spec.log 2>&1; tail -1 gpurun_out/r2_ncu_nd_spec.log
