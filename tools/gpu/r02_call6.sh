#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== stream tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x -k "two_stream or wgrad_side or plain_vae_step_vs_golden" > gpurun_out/r02_stream_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_stream_tests.log
echo "== A/B two streams"; bash tools/gpu/ab.sh SIVAE_TWO_STREAMS 0 1
echo "== A/B wgrad stream"; bash tools/gpu/ab.sh SIVAE_WGRAD_STREAM 0 1
echo "== both"; for i in 1 2; do SIVAE_TWO_STREAMS=1 SIVAE_WGRAD_STREAM=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-lshape 2>gpurun_out/both.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('both', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'vol/s', d['config'].get('cuda_graph'), d['config'].get('cuda_graph_note'))"; done; tail -3 gpurun_out/both.err
bash tools/gpu/ncu_profile.sh
