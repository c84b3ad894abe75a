#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== N=2 graph"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-400
echo "== N=2 reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2_ref.log | cut -c1-200
echo "== N=1"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-300
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
