#!/bin/bash
# N-GPU pass (N = $1, default 2): sharded ADMM check (NCCL), then bench at N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_admm_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist_check_n$N.log
grep -E "rank|rc=|Error|error" gpurun_out/dist_check_n$N.log | tail -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_n$N.log
tail -2 gpurun_out/bench_n$N.log | cut -c1-300
