#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv3_|to1|c1_to|wgrad_c1_tc_kernel|bn_stats_kernel|bn_act_fwd|bn_bwd_.*plain' -s 14 -c 16 \
    -o gpurun_out/r01c_hot python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_targets.log; ls -la gpurun_out/*.ncu-rep
