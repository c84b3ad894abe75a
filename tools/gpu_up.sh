#!/bin/bash
mkdir -p gpurun_out
echo "== fused"; timeout 120 python tools/upconv_bench.py 2>&1 | tail -3
echo "== tapwise"; SIVAE_UPCONV_FUSED=0 timeout 120 python tools/upconv_bench.py 2>&1 | tail -3
timeout 120 python tools/upconv_bench.py 8 40 48 40 64 64 1 > gpurun_out/up_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upconv3_fused' -c 2 -o gpurun_out/r01c_up python tools/upconv_bench.py 8 40 48 40 64 64 1 > gpurun_out/ncu_up.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_up.log
