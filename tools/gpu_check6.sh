#!/bin/bash
mkdir -p gpurun_out
echo "== conv"; timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_conv.log
echo "== convbench kwcopy"; timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 3 2>&1 | tail -3
echo "== convbench tapwise"; SIVAE_CONV_TAPWISE=1 timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 3 2>&1 | tail -3
echo "== convbench 40"; timeout 120 python tools/conv_bench.py 8 40 48 40 64 64 3 2>&1 | tail -3
echo "== model"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_model.log
echo "== bench graph"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-260
