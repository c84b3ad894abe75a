#!/bin/bash
# Runs on the GPU box (via gpurun): probe + parity tests + a short bench, logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe"; timeout 600 python tools/conv_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -40 gpurun_out/probe.log
echo "== kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_kernels.log
echo "== conv"; timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_conv.log
echo "== model"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t_model.log
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench.log
