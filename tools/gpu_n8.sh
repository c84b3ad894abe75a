#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
echo "== N=$NG graph"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n8.log | cut -c1-420
