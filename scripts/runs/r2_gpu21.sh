#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "rotated or adjoint or multi" > gpurun_out/r2_pytest_cell3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_cell3.log
tail -4 gpurun_out/r2_pytest_cell3.log
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 > gpurun_out/r2_cg_rigid_cellv3.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_cellv3.log
NOPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rot_adjoint_cell_kernel|rot_forward_kernel" -s 10 -c 2 -o gpurun_out/r2_prof_rotcell3 python scripts/microbench_cg.py sr3_256_rigid 20 1 > gpurun_out/r2_ncu_rotcell3.log 2>&1; tail -1 gpurun_out/r2_ncu_rotcell3.log
