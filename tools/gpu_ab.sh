#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
echo "== cluster BN on"; timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])"
echo "== cluster BN off"; SIVAE_NO_BN_CLUSTER=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])"
done
