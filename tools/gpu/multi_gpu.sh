#!/bin/bash
# Weak-scaling check on every GPU of the box (gpurun --gpus N -- bash tools/gpu/multi_gpu.sh): our arm as three CUDA
# graphs + two NCCL all-reduces per step, then the reference arm (rank 0 only works), launched as the driver does.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
echo "== N=$NG"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n$NG.log | cut -c1-420
echo "== N=$NG reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 > gpurun_out/bench_n${NG}_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n${NG}_ref.log | cut -c1-200
