#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
lscpu | grep -i -E "numa|^CPU\(s\)|model name" >> gpurun_out/r2_topo_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_admm_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; tail -$N gpurun_out/r2_dist_check_n$N.log | cut -c1-160
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_n$N.log
tail -2 gpurun_out/r2_bench_n$N.log | cut -c1-200
