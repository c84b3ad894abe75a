#!/bin/bash
# A/B/A/B of two builds of the library on the headline bench:  ab_lib.sh <libA> <libB> [pytest -k expression]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [ -n "$3" ]; then timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$3" 2>&1 | tail -3; fi
for i in 1 2; do
for lib in "$1" "$2"; do
  SIVAE_LIB=$PWD/$lib timeout 600 python bench.py --steps 8 --warmup 3 --no-lshape --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'vol/s', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done; done
