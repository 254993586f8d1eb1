#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-sharded --steps 5 --warmup 3 > gpurun_out/r2_bench_steady.log 2>&1
grep '^{' gpurun_out/r2_bench_steady.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms %.3f e2e %.0f energy %.0f roofline %.3f launch %.2f us' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['energy_rule']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']*1e3))"
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu5.log
tail -3 gpurun_out/r2_pytest_gpu5.log
