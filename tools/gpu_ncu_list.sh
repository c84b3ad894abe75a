#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
