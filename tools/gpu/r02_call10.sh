#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== lshape probe rand"; timeout 600 python tools/lshape_probe.py 160 192 160 2 12 rand > gpurun_out/r02_lshape_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_lshape_probe.log | tail -22
echo "== lshape probe blobs"; timeout 600 python tools/lshape_probe.py 160 192 160 2 12 blobs > gpurun_out/r02_lshape_probe_blobs.log 2>&1; echo "rc=$?"; grep ours gpurun_out/r02_lshape_probe_blobs.log | tail -12
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -8 gpurun_out/kernel_table.txt
