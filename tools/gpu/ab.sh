#!/bin/bash
# A/B/A/B of one environment knob on one box:  bash tools/gpu/ab.sh SIVAE_N256 0 4   (optional 4th arg: extra bench flags)
# Knobs: SIVAE_CONV_KD, SIVAE_CONV_KW, SIVAE_N256, SIVAE_DEEP_RING, SIVAE_UPCONV_FUSED, SIVAE_WGRAD_KW, SIVAE_UPWGRAD_TALL,
#        SIVAE_BN_CLUSTER, SIVAE_NO_FUSED_STATS.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VAR=$1; A=$2; B=$3; EXTRA=$4
for i in 1 2; do
  for v in "$A" "$B"; do
    env $VAR=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-lshape $EXTRA 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'vol/s', d['gpu_launches'], 'launches')"
  done
done
