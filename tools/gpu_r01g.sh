#!/bin/bash
mkdir -p gpurun_out
echo "== kernels tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short -x > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_kernels.log
echo "== model"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; grep -E "rel |passed|failed|Error|cosine" gpurun_out/t_model.log | tail -16
echo "== pointwise"; timeout 300 python tools/pointwise_bench.py > gpurun_out/pointwise.log 2>&1; cat gpurun_out/pointwise.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300; head -6 gpurun_out/kernel_table.txt
