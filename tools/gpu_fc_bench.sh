#!/bin/bash
# FC-latent variant (BASELINE config 2): bench line + per-kernel table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --workload fc600 --batch 4 --steps 5 --warmup 3 --kernel-table gpurun_out/fc600_kernel_table.txt \
  > gpurun_out/fc600_bench.json 2> gpurun_out/fc600_bench.err
echo "fc600 bench rc=$?"
tail -c 1500 gpurun_out/fc600_bench.json
tail -3 gpurun_out/fc600_bench.err
head -30 gpurun_out/fc600_kernel_table.txt
