#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; tail -5 gpurun_out/bench_default.log | cut -c1-2500
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; tail -5 gpurun_out/bench_reference.log | cut -c1-900
nproc; python -c "import torch; print(torch.get_num_threads())"
