#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== determinism 40x48x40"; timeout 600 python tools/determinism_probe.py --vol 40 48 40 --batch 4 --repeats 4 > gpurun_out/r02_determinism.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02_determinism.log
echo "== determinism 80x96x80 b8 philox"; timeout 600 python tools/determinism_probe.py --vol 80 96 80 --batch 8 --repeats 3 --philox >> gpurun_out/r02_determinism.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02_determinism.log
echo "== bisect following ours"; timeout 900 python tools/grad_bisect.py --vol 40 48 40 --batch 4 --at 60 120 199 --follow ours --worst 8 > gpurun_out/r02_grad_bisect_ours.log 2>&1; echo "rc=$?"; grep -E "step|mean cos|^terms|^oracle |^ours/|^control " gpurun_out/r02_grad_bisect_ours.log
echo "== new parity tests"; timeout 1200 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s -k "plain_vae or config1 or bench_config" > gpurun_out/r02_newtests.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/r02_newtests.log | tail -45
