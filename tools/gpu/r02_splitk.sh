#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== conv + fc + model tests"; timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_fc_gpu.py tests/test_guard_bands_gpu.py -q -m gpu --tb=short > gpurun_out/r02_splitk_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_splitk_tests.log
echo "== A/B fc600"; bash tools/gpu/ab.sh SIVAE_SPLITK 0 1 "--workload fc600 --batch 4"
echo "== headline unaffected"; bash tools/gpu/ab.sh SIVAE_SPLITK 0 1
