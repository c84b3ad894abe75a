#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_kernels.log
echo "== model"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t_model.log
echo "== bench graph"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
echo "== bench eager"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_eager.log | cut -c1-400
