#!/bin/bash
# Profiles for profiles/ (one GPU; each ncu pass only after the same command exited 0 without ncu):
#   1. launch list of one eager training step (per-kernel shares, tools/launch_summary.py reads it)
#   2. ncu --set full of the hot kernels (tools/ncu_targets.py launches each once) -> tools/ncu_summarize.py (run in the
#      authoring container on the pulled .ncu-rep) writes profiles/<round>_ncu_hot_kernels.md + profiles/top_kernel_ncu.json
# bench.py runs with --no-clocks here: the nvidia-smi sampler child is what made the round-1 launch-list pass hang under ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-graph --no-clocks --no-lshape"
echo "== ncu launch list"
timeout 300 $BENCH > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2900 --csv --log-file gpurun_out/launches.csv \
    $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full, hot kernels"
timeout 300 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && \
timeout 1500 ncu --target-processes application-only --set full --clock-control none --import-source on \
    -k regex:'conv3_kd3|conv3_wgrad_kw64|upconv3_fused|upconv3_wgrad_tall|conv3_to1_halo|c1_to_c64_tc_kernel' -s 6 -c 6 \
    -o gpurun_out/hot -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_targets.log
