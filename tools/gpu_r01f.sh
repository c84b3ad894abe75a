#!/bin/bash
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short -x > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_conv.log
echo "== convbench auto (kd fprop, kw wgrad)"; timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 5 2>&1 | tail -3
echo "== convbench tapwise"; SIVAE_CONV_KD=0 SIVAE_WGRAD_KW=0 timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 5 2>&1 | tail -3
echo "== convbench 40 auto"; timeout 120 python tools/conv_bench.py 8 40 48 40 64 64 5 2>&1 | tail -3
echo "== convbench 40 tapwise"; SIVAE_CONV_KD=0 SIVAE_WGRAD_KW=0 timeout 120 python tools/conv_bench.py 8 40 48 40 64 64 5 2>&1 | tail -3
echo "== convbench 40 64->128 auto"; timeout 120 python tools/conv_bench.py 8 40 48 40 64 128 5 2>&1 | tail -3
echo "== convbench 40 64->128 generic wgrad"; SIVAE_WGRAD_KW=0 timeout 120 python tools/conv_bench.py 8 40 48 40 64 128 5 2>&1 | tail -3
echo "== model"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; grep -E "rel |passed|failed|Error|cosine" gpurun_out/t_model.log | tail -24
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -12 gpurun_out/kernel_table.txt
