#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -k "tail_dgrad or fused_decoder_end" > gpurun_out/r02_tail_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_tail_tests.log
echo "== micro bench"; timeout 300 python tools/tail_bench.py
echo "== ncu"; timeout 600 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:'c1_to_c64_tc_kernel' -s 40 -c 2 -o gpurun_out/tail -f python tools/tail_bench.py > gpurun_out/ncu_tail.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tail.log
echo "== A/B"; bash tools/gpu/ab.sh SIVAE_FUSE_TAIL 0 1
