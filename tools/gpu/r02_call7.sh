#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== all gpu tests (streams on)"; SIVAE_TWO_STREAMS=1 SIVAE_WGRAD_STREAM=1 timeout 2400 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all_streams.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_all_streams.log
bash tools/gpu/ncu_profile.sh
