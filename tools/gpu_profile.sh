#!/bin/bash
# GPU box: parity re-check, plain bench, then ncu launch list of one bench step and a full capture of the conv kernels.
mkdir -p gpurun_out
echo "== conv"; timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_conv.log
echo "== model"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/t_model.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_model.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
echo "== ncu launches"
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
echo "== ncu full conv"
timeout 300 python tools/conv_bench.py 8 80 96 80 64 64 2 > gpurun_out/conv_bench_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_ -c 4 -o gpurun_out/conv_prof \
    python tools/conv_bench.py 8 80 96 80 64 64 2 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; cat gpurun_out/conv_bench_plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
