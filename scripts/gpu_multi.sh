#!/bin/bash
# 2-GPU pass: sharded ADMM check, then bench at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_admm_check.py > gpurun_out/dist_check.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist_check.log
grep -E "rank|rc=|Error|error" gpurun_out/dist_check.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_n2.log
tail -2 gpurun_out/bench_n2.log | cut -c1-400
