#!/bin/bash
mkdir -p gpurun_out
echo "== bench (with cpu baseline)"; timeout 900 python bench.py --steps 5 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
echo "== ncu launches"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
