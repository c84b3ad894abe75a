#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== A/B pointwise grid cap"; bash tools/gpu/ab.sh SIVAE_PW_BLOCKS 1184 2368
echo "== third value"; for v in 592 4736; do env SIVAE_PW_BLOCKS=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-lshape 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SIVAE_PW_BLOCKS=$v', round(d['ms_per_step'],2), 'ms')"; done
