#!/bin/bash
# thin-conv kernel changes: parity tests of the kernels + model, thin-kernel bandwidths, headline bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_guard_bands_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== pointwise"
timeout 300 python tools/pointwise_bench.py 2>&1 | tee gpurun_out/pointwise.txt | grep -i "c1\|taps"
echo "== bench"
timeout 600 python bench.py --steps 8 --warmup 3 --no-lshape --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['loss'])
PY
grep -i "to1\|c1" gpurun_out/kernel_table.txt | head
