#!/bin/bash
# L2 hints under the real bench configuration (3 channel streams) and at 384^3
mkdir -p gpurun_out
for h in 0 1 4 5; do
  timeout 300 python bench.py --no-cpu-baseline --no-sharded --steps 5 --warmup 3 --tune l2_hints=$h > gpurun_out/r2_bench_hints$h.log 2>&1
  tail -1 gpurun_out/r2_bench_hints$h.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hints $h value %.0f ms %.3f e2e %.0f energy %.0f roofline %.3f launch %.2f us' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['energy_rule']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']*1e3))"
done
for h in 0 1 4 5; do
  NOPROF=1 timeout 300 python scripts/microbench_cg.py thickz2_384 20 3 l2_hints=$h 2>&1 | tail -1 | cut -c1-60
done
