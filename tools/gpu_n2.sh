#!/bin/bash
# 2-GPU check of the multi-rank path: graph mode (3 graphs + NCCL between replays) and eager bucketed reducer.
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== N=2 graph"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_n2.log | cut -c1-700
echo "== N=2 eager"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-graph > gpurun_out/bench_n2_eager.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2_eager.log | cut -c1-300
echo "== N=1"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-300
