#!/bin/bash
# compute-sanitizer on the cell adjoint (memcheck + racecheck: the colour passes are plain RMW on
# shared memory), and the fit tests after the tolerance note
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "cell_adjoint" > gpurun_out/r2_sanitizer_cell_mem.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_cell_mem.log
tail -5 gpurun_out/r2_sanitizer_cell_mem.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "cell_adjoint_any_rotation or (cell_adjoint_equals and sr2)" > gpurun_out/r2_sanitizer_cell_race.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_cell_race.log
tail -5 gpurun_out/r2_sanitizer_cell_race.log
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q > gpurun_out/r2_pytest_fit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_fit.log
tail -3 gpurun_out/r2_pytest_fit.log
