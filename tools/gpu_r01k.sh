#!/bin/bash
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short -x -k "upconv" > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_conv.log
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_all.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -12 gpurun_out/kernel_table.txt
