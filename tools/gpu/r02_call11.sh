#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== guard bands"; timeout 1200 python -m pytest tests/test_guard_bands_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -k "guard or bounds or projection" > gpurun_out/r02_guard.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_guard.log
echo "== bench"; timeout 900 python bench.py --steps 8 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; tail -2 gpurun_out/bench.err
echo "== fc600"; timeout 600 python bench.py --workload fc600 --batch 4 --steps 8 --warmup 3 --kernel-table gpurun_out/fc600_kernel_table.txt > gpurun_out/fc600_bench.json 2> gpurun_out/fc600_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/fc600_bench.json
