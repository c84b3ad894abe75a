#!/bin/bash
mkdir -p gpurun_out
echo "== bench L-shape"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --vol 160 192 160 --batch 2 > gpurun_out/bench_L.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_L.log | cut -c1-400
echo "== ncu launches"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
