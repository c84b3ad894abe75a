#!/bin/bash
# deep-ring igemm A/B + linear roofline + fc tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fprop or dgrad" > gpurun_out/conv_tests.log 2>&1
echo "conv tests rc=$?"; tail -2 gpurun_out/conv_tests.log
timeout 300 python tools/linear_bench.py > gpurun_out/linear_bench.md 2>&1; cat gpurun_out/linear_bench.md
for ring in 0 auto; do
  if [ $ring = auto ]; then unset SIVAE_DEEP_RING; else export SIVAE_DEEP_RING=$ring; fi
  timeout 600 python bench.py --workload fc600 --batch 4 --steps 5 --warmup 3 --kernel-table gpurun_out/fc600_kernel_table_$ring.txt \
    > gpurun_out/fc600_bench_$ring.json 2> gpurun_out/fc600_bench_$ring.err
  echo "fc600 ring=$ring rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/fc600_bench_$ring.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
  grep "5, 6, 5, 256, 256\|10, 12, 10, 128, 128" gpurun_out/fc600_kernel_table_$ring.txt
done
unset SIVAE_DEEP_RING
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_deep.json 2> gpurun_out/bench_deep.err
python -c "
import json; d=json.load(open('gpurun_out/bench_deep.json')); print('headline', d['value'], d['ms_per_step'], d['e2e']['value'])"
