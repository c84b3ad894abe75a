#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_kernels.log
echo "== conv"; timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_conv.log
echo "== model"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_model.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
echo "== ncu launches"
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
