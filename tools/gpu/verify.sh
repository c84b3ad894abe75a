#!/bin/bash
# Round-end verification on one B200 (run through gpurun):  bash tools/gpu/verify.sh
# all GPU tests, __graft_entry__.smoke(), headline bench (cpu baseline, L-shape entry, per-kernel table), FC-latent bench,
# reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 2400 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400; head -12 gpurun_out/kernel_table.txt
echo "== fc600"; timeout 600 python bench.py --workload fc600 --batch 4 --steps 5 --warmup 3 --kernel-table gpurun_out/fc600_kernel_table.txt > gpurun_out/fc600_bench.json 2> gpurun_out/fc600_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/fc600_bench.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
