#!/bin/bash
mkdir -p gpurun_out
echo "== tests retrieval/optim"; timeout 900 python -m pytest tests/test_retrieval.py tests/test_optim.py -q -m gpu --tb=short > gpurun_out/t_new.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_new.log
echo "== latent bench"; timeout 600 python tools/latent_bench.py > gpurun_out/latent_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/latent_bench.log
