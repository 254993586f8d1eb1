#!/bin/bash
# ur_backproject (one-pass initial estimate of the e2e pipeline): parity + bench e2e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q -x -k "backproject or pipeline" > gpurun_out/r2_pytest_bp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_bp.log
tail -5 gpurun_out/r2_pytest_bp.log
timeout 300 python bench.py --no-cpu-baseline --no-sharded --steps 5 --warmup 3 > gpurun_out/r2_bench_bp.log 2>&1
grep '^{' gpurun_out/r2_bench_bp.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms %.3f e2e %.0f (launches %d) energy %.0f roofline %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['gpu_launches'], d['energy_rule']['value'], d['roofline']['frac']))"
tail -3 gpurun_out/r2_bench_bp.log | cut -c1-200
