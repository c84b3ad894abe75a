#!/bin/bash
# Round 2, call 1: bisect the 40x48x40 loss-curve divergence + first run of the split-K tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== grad bisect"; timeout 900 python tools/grad_bisect.py --vol 40 48 40 --batch 4 --at 0 40 120 > gpurun_out/r02_grad_bisect.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_grad_bisect.log
echo "== loss curve, generic kernels"; timeout 900 python tests/loss_curve.py --steps 200 --vol 40 48 40 --batch 4 --control --out gpurun_out/r02_lc_generic --env SIVAE_CONV_KD=0 SIVAE_UPCONV_FUSED=0 SIVAE_WGRAD_KW=0 SIVAE_UPWGRAD_TALL=0 SIVAE_NO_FUSED_STATS=1 SIVAE_TO1_TAPWISE=1 > gpurun_out/r02_lc_generic.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/r02_lc_generic.log
echo "== splitk tests"; SIVAE_TEST_SPLITK=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -k splitk --tb=short > gpurun_out/r02_splitk.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_splitk.log
