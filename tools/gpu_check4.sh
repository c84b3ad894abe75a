#!/bin/bash
mkdir -p gpurun_out
echo "== kernels+conv"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_kernels.log
echo "== model"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_model.log
echo "== bench graph"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench.log
echo "== ncu launches"
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches.csv
