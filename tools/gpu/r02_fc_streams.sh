#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fc + model + loss-curve tests"; timeout 1800 python -m pytest tests/test_fc_gpu.py tests/test_model_gpu.py tests/test_loss_curve.py -q -m gpu --tb=short > gpurun_out/r02_fc_stream_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_fc_stream_tests.log
echo "== A/B fc600"; bash tools/gpu/ab.sh SIVAE_TWO_STREAMS 0 1 "--workload fc600 --batch 4"
