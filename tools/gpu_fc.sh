#!/bin/bash
# FC-latent variant: kernel + model parity on the GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fc_gpu.py -q -m gpu -x -s > gpurun_out/fc_tests.log 2>&1
echo "fc tests rc=$?"
tail -5 gpurun_out/fc_tests.log
