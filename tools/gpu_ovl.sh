#!/bin/bash
# dgrad/wgrad two-stream overlap for small layers: parity + A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fc_gpu.py -q -m gpu -x > gpurun_out/ovl_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/ovl_tests.log
for ovl in 0 1 0 1; do
  SIVAE_BWD_OVERLAP=$ovl timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_$ovl.json 2> gpurun_out/b_$ovl.err
  SIVAE_BWD_OVERLAP=$ovl timeout 600 python bench.py --workload fc600 --batch 4 --steps 5 --warmup 3 > gpurun_out/f_$ovl.json 2> gpurun_out/f_$ovl.err
  python -c "
import json
d=json.load(open('gpurun_out/b_$ovl.json')); f=json.load(open('gpurun_out/f_$ovl.json'))
print('overlap=$ovl headline', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'graph', d['config']['cuda_graph'], '| fc600', round(f['ms_per_step'],2), 'ms', round(f['value'],1), 'graph', f['config']['cuda_graph'])"
done
