#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== new parity tests"; timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_optim.py -q -m gpu --tb=short -s -k "plain_vae or config1 or bench_config or resume" > gpurun_out/r02_newtests.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/r02_newtests.log | tail -60
echo "== ensemble default"; timeout 1200 python tools/curve_ensemble.py --vol 40 48 40 --batch 4 --steps 200 --replicas 4 --out gpurun_out/r02_ensemble_default > gpurun_out/r02_ensemble_default.log 2>&1; echo "rc=$?"; tail -75 gpurun_out/r02_ensemble_default.log
echo "== ensemble generic"; timeout 1200 python tools/curve_ensemble.py --vol 40 48 40 --batch 4 --steps 200 --replicas 3 --out gpurun_out/r02_ensemble_generic --env SIVAE_CONV_KD=0 SIVAE_UPCONV_FUSED=0 SIVAE_WGRAD_KW=0 SIVAE_UPWGRAD_TALL=0 SIVAE_NO_FUSED_STATS=1 SIVAE_TO1_TAPWISE=1 > gpurun_out/r02_ensemble_generic.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02_ensemble_generic.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_all.log
