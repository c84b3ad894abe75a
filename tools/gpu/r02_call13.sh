#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fused tail kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_guard_bands_gpu.py -q -m gpu --tb=short -k "tail_dgrad or thin_convs or c1_to_cn" > gpurun_out/r02_tail_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_tail_tests.log
echo "== fused decoder end model test"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -k "fused_decoder_end" > gpurun_out/r02_tail_model.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_tail_model.log
echo "== A/B"; bash tools/gpu/ab.sh SIVAE_FUSE_TAIL 0 1
