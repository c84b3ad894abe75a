#!/bin/bash
# ncu --set full of the memory-bound kernels (tools/ncu_targets_pointwise.py launches each once at the full-resolution
# size): DRAM bytes and throughput per launch.  The raw page is exported on the box (the report itself can exceed what
# gpurun copies back); tools/ncu_summarize.py --csv turns it into profiles/<round>_ncu_pointwise.md.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/ncu_targets_pointwise.py > gpurun_out/pointwise_targets_plain.txt 2>&1 && \
timeout 900 ncu --target-processes application-only --set full --clock-control none \
    -k regex:'bn_act_plain_fwd|bn_act_fwd_kernel|bn_bwd_reduce_plain|bn_bwd_apply_plain|bn_act_bwd_reduce|bn_act_bwd_apply|mse_persample' \
    -c 12 -o /tmp/pointwise -f python tools/ncu_targets_pointwise.py > gpurun_out/ncu_pointwise.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_pointwise.log
ncu -i /tmp/pointwise.ncu-rep --page raw --csv > gpurun_out/pointwise_raw.csv 2>/dev/null; ls -la /tmp/pointwise.ncu-rep gpurun_out/pointwise_raw.csv
rm -f gpurun_out/pointwise.ncu-rep
