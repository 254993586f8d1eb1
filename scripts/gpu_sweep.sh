#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream_kernel.py tests/test_gpu_solver.py -m gpu -q -x 2>&1 | tail -3
(for rr in 0 1; do echo "== r_reverse $rr"; NOPROF=1 timeout 120 python scripts/microbench_cg.py sr3_256 20 5 r_reverse=$rr 2>&1 | tail -3 | cut -c1-60; done
for rr in 0 1; do echo "== r_reverse $rr thickz2"; NOPROF=1 timeout 120 python scripts/microbench_cg.py thickz2_256 20 5 r_reverse=$rr 2>&1 | tail -3 | cut -c1-60; done
) 2>&1 | tee gpurun_out/sweep4.log
