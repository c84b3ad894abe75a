#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fc_gpu.py -q -m gpu -x -k "linear or layout" > gpurun_out/lin_tests.log 2>&1
echo "linear tests rc=$?"; tail -2 gpurun_out/lin_tests.log
timeout 300 python tools/linear_bench.py > gpurun_out/linear_bench.md 2>&1; cat gpurun_out/linear_bench.md
