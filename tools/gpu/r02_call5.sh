#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== warm ensemble 40x48x40"; timeout 1500 python tools/curve_ensemble.py --vol 40 48 40 --batch 4 --steps 200 --replicas 4 --warm-start 20 --out gpurun_out/r02_ensemble_warm > gpurun_out/r02_ensemble_warm.log 2>&1; echo "rc=$?"; head -5 gpurun_out/r02_ensemble_warm.log; tail -8 gpurun_out/r02_ensemble_warm.log
echo "== warm curve 80x96x80 b4"; timeout 2400 python tests/loss_curve.py --steps 200 --vol 80 96 80 --batch 4 --control --warm-start 20 --out gpurun_out/r02_loss_curve_80x96x80_warm > gpurun_out/r02_lc_80_warm.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/r02_lc_80_warm.log
echo "== tests"; timeout 1200 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s -k "plain_vae or config1" > gpurun_out/r02_newtests2.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/r02_newtests2.log | tail -12
