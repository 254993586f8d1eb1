#!/bin/bash
# 8-GPU pass: bench only (weak scaling: one 3-channel subject per GPU)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_n8.log
tail -2 gpurun_out/bench_n8.log | cut -c1-300
