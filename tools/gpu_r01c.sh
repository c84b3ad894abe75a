#!/bin/bash
# GPU box: full parity suite, bench (with cpu baseline), launch list of one eager step.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
echo "== ncu launches"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
