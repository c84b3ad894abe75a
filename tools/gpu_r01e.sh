#!/bin/bash
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short -x > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_conv.log
echo "== convbench kw"; timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 5 2>&1 | tail -3
echo "== convbench tapwise"; SIVAE_CONV_KW=0 timeout 120 python tools/conv_bench.py 8 80 96 80 64 64 5 2>&1 | tail -3
echo "== convbench 40 kw"; timeout 120 python tools/conv_bench.py 8 40 48 40 64 64 5 2>&1 | tail -3
echo "== convbench 40 tapwise"; SIVAE_CONV_KW=0 timeout 120 python tools/conv_bench.py 8 40 48 40 64 64 5 2>&1 | tail -3
echo "== convbench 40 64->128 kw"; timeout 120 python tools/conv_bench.py 8 40 48 40 64 128 5 2>&1 | tail -3
echo "== convbench 40 64->128 tapwise"; SIVAE_CONV_KW=0 timeout 120 python tools/conv_bench.py 8 40 48 40 64 128 5 2>&1 | tail -3
echo "== model"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_model.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -8 gpurun_out/kernel_table.txt
echo "== loss curve small + control"; timeout 900 python tools/loss_curve.py --steps 200 --vol 16 24 16 --batch 2 --control --out gpurun_out/loss_curve_small > gpurun_out/loss_curve_small.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/loss_curve_small.log
echo "== loss curve 40x48x40 + control"; timeout 1200 python tools/loss_curve.py --steps 200 --vol 40 48 40 --batch 4 --control --out gpurun_out/loss_curve > gpurun_out/loss_curve.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/loss_curve.log
