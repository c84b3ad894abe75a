#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "bn" 2>&1 | tail -3
timeout 300 python tools/pointwise_bench.py 2>&1 | grep "bn_"
bash tools/gpu/r02_ab_lib.sh build_old/libsivae_old.so soft-intro-vae-for-3d-mri_b200/libsivae.so
