#!/bin/bash
# lean kernel: tile rows / segments / stencil lag sweep
mkdir -p gpurun_out
timeout 500 python scripts/r2_sweep_tile.py sr3_256 > gpurun_out/r2_sweep_tile.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sweep_tile.log
tail -3 gpurun_out/r2_sweep_tile.log
