#!/bin/bash
# multi-GPU pass: N = number of GPUs of this box
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
numactl -H >> gpurun_out/r2_topo_n$N.txt 2>&1 || lscpu | grep -i numa >> gpurun_out/r2_topo_n$N.txt
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -q > gpurun_out/r2_pytest_multidev_n$N.log 2>&1; tail -3 gpurun_out/r2_pytest_multidev_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_admm_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; tail -$N gpurun_out/r2_dist_check_n$N.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_n$N.log
tail -2 gpurun_out/r2_bench_n$N.log | cut -c1-3000
