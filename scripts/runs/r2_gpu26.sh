#!/bin/bash
mkdir -p gpurun_out
for shape in 0 1 2; do
NOPROF=1 timeout 300 python scripts/microbench_cg.py sr3_256_rigid 20 3 rot_cell_shape=$shape > gpurun_out/r2_cg_rigid_shape$shape.log 2>&1; tail -3 gpurun_out/r2_cg_rigid_shape$shape.log
done
cat > /tmp/shape_test.py <<'PY'
import sys, pytest
from unires_b200 import _lib
_lib.check(_lib.lib.ur_tune(b'rot_cell_shape', int(sys.argv[1])))
sys.exit(pytest.main(['tests/test_gpu_ops.py', '-m', 'gpu', '-q', '-x', '-k', 'rotated or adjoint']))
PY
for shape in 1 2; do
timeout 600 python /tmp/shape_test.py $shape > gpurun_out/r2_pytest_shape$shape.log 2>&1; tail -2 gpurun_out/r2_pytest_shape$shape.log
done
