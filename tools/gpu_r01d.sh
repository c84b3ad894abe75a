#!/bin/bash
mkdir -p gpurun_out
echo "== loss curve small"; timeout 900 python tools/loss_curve.py --steps 200 --vol 16 24 16 --batch 2 --out gpurun_out/loss_curve_small > gpurun_out/loss_curve_small.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/loss_curve_small.log
echo "== loss curve 40x48x40"; timeout 1200 python tools/loss_curve.py --steps 200 --vol 40 48 40 --batch 4 --out gpurun_out/loss_curve > gpurun_out/loss_curve.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/loss_curve.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; cat gpurun_out/kernel_table.txt
