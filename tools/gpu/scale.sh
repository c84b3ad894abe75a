#!/bin/bash
# Weak scaling on one box, launched as the driver launches it:  gpurun --gpus 8 -- bash tools/gpu/scale.sh "1 8"
# (every N of the list back to back; the box time is charged x the GPUs of the box, so keep the list short)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for N in ${1:-1 8}; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-lshape > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + N)) \
      bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-lshape > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "rc=$?"
  python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/scale_n{n}.json').read().strip().splitlines()[-1])
    print(f"N={n}", round(d['ms_per_step'], 2), 'ms', round(d['value'], 1), 'vol/s', 'e2e', round(d['e2e']['value'], 1),
          'params_identical', d.get('params_identical'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(f"N={n} failed:", e)
    print(open(f'gpurun_out/scale_n{n}.err').read()[-1500:])
PY
done
