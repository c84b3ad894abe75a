#!/bin/bash
mkdir -p gpurun_out
echo "== bench (with cpu baseline)"; timeout 900 python bench.py --steps 5 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
echo "== ncu launches"
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3100 -c 1700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full hot kernels"
timeout 300 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv3_kd3|conv3_wgrad_kw64|upconv3_fused|upconv3_wgrad_tall|conv3_to1_halo|c1_to_c64_tc_kernel' -s 7 -c 7 \
    -o gpurun_out/r01e_hot python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_targets.log
