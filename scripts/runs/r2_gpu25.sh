#!/bin/bash
# fused cell kernel (adjoint + D'D + epilogue in one launch): parity, timing, tile-shape test of the lean kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_midsize.py -m gpu -q -x > gpurun_out/r2_pytest_cell5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell5.log
tail -4 gpurun_out/r2_pytest_cell5.log
timeout 900 python -m pytest tests/test_gpu_stream_kernel.py -m gpu -q -x -k "tile_rows" > gpurun_out/r2_pytest_tiles.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_tiles.log
tail -4 gpurun_out/r2_pytest_tiles.log
timeout 900 python -m pytest tests/test_gpu_fullsize_oracle.py -m gpu -q -x -k "rigid" > gpurun_out/r2_pytest_cell5_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell5_full.log
tail -4 gpurun_out/r2_pytest_cell5_full.log
for cell in 2 1; do
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 rot_cell=$cell > gpurun_out/r2_cg_rigid_cellv5_$cell.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_cellv5_$cell.log
done
NOPROF=1 timeout 300 python scripts/microbench_admm.py sr3_256_rigid > gpurun_out/r2_admm_rigid_v5.log 2>&1; tail -5 gpurun_out/r2_admm_rigid_v5.log
