#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_all.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -14 gpurun_out/kernel_table.txt
