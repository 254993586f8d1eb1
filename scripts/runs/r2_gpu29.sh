#!/bin/bash
# steady / edge trip split of the lean kernel: parity subset, timing, instruction count
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_kernel.py -m gpu -q -x > gpurun_out/r2_pytest_steady.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_steady.log
tail -3 gpurun_out/r2_pytest_steady.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256 20 5 > gpurun_out/r2_cg_sr3_steady.log 2>&1; tail -3 gpurun_out/r2_cg_sr3_steady.log | cut -c1-60
NOPROF=1 timeout 300 python scripts/microbench_cg.py thickz2_384 20 3 2>&1 | tail -1 | cut -c1-60
NOPROF=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"lhs_fast_kernel" -s 8 -c 1 --csv --log-file gpurun_out/r2_steady_inst.csv python scripts/microbench_cg.py sr3_256 20 1 > /dev/null 2>&1; grep -v "^==" gpurun_out/r2_steady_inst.csv | awk -F'","' '{print $13, $15}' | cut -c1-150
