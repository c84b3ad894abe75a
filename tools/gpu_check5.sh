#!/bin/bash
mkdir -p gpurun_out
echo "== kernels+conv"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short -x > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_kernels.log
echo "== model"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; grep -E "rel |passed|failed|Error" gpurun_out/t_model.log | tail -20
echo "== pointwise"; timeout 300 python tools/pointwise_bench.py > gpurun_out/pointwise.log 2>&1; cat gpurun_out/pointwise.log
echo "== bench graph"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
