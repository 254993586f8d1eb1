#!/bin/bash
mkdir -p gpurun_out
(for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120; done
) | tee gpurun_out/variance2.log
