#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short > gpurun_out/r02_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_kernels.log
echo "== layers headline"; timeout 600 python tools/parity_probe.py layers --vol 40 48 40 --batch 4 --train-steps 0 40 > gpurun_out/r02_layers_headline.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r02_layers_headline.log
echo "== layers config1"; timeout 600 python tools/parity_probe.py layers --net config1 --vol 80 96 80 --batch 2 --train-steps 0 > gpurun_out/r02_layers_config1.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/r02_layers_config1.log
echo "== bias"; timeout 900 python tools/parity_probe.py bias --vol 40 48 40 --batch 4 --train-steps 40 --draws 12 > gpurun_out/r02_bias40.log 2>&1; echo "rc=$?"; tail -70 gpurun_out/r02_bias40.log
echo "== A/B splitk"; bash tools/gpu/ab.sh SIVAE_SPLITK 0 1
echo "== A/B keep bits"; bash tools/gpu/ab.sh SIVAE_KEEP_BITS 0 1
