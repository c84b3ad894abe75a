#!/bin/bash
# 2 GPUs (gpurun --gpus 2 -- bash tools/gpu/two_gpu.sh): NCCL equivalence test, N=1 vs N=2 bench (three graphs + NCCL between
# replays vs NCCL captured in one graph), --global-batch 64
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== nccl equivalence test"; timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -s --tb=short > gpurun_out/r02_multi_gpu_test.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_multi_gpu_test.log
echo "== N=1"; timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-lshape > gpurun_out/r02_n1.json 2> gpurun_out/r02_n1.err; echo "rc=$?"; cut -c1-260 gpurun_out/r02_n1.json
echo "== N=2 three graphs"; timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_n2.json 2> gpurun_out/r02_n2.err; echo "rc=$?"; cut -c1-260 gpurun_out/r02_n2.json; grep -o '"params_identical": [a-z]*' gpurun_out/r02_n2.json
echo "== N=2 NCCL inside one graph"; SIVAE_GRAPH_NCCL=1 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_n2_graphnccl.json 2> gpurun_out/r02_n2_graphnccl.err; echo "rc=$?"; cut -c1-260 gpurun_out/r02_n2_graphnccl.json; grep -o '"params_identical": [a-z]*\|"cuda_graph_note": [^,]*' gpurun_out/r02_n2_graphnccl.json; tail -3 gpurun_out/r02_n2_graphnccl.err
echo "== N=2 global batch 64 (local 32)"; timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --global-batch 64 > gpurun_out/r02_n2_gb64.json 2> gpurun_out/r02_n2_gb64.err; echo "rc=$?"; cut -c1-260 gpurun_out/r02_n2_gb64.json; tail -2 gpurun_out/r02_n2_gb64.err
