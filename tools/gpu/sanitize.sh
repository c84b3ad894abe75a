#!/bin/bash
# NOTE: this GPU pool refuses compute-sanitizer ("closed on this pool": runs under it have left GPUs needing a reset), so this
# script has not produced a log here; tests/test_guard_bands_gpu.py and tools/determinism_probe.py stand in for it.
# compute-sanitizer over the small-shape kernel parity tests: memcheck on every kernel family, racecheck
# on the hand-rolled mbarrier / TMA / TMEM pipelines.  Each pass under its own timeout (the tools slow kernels 10-100x);
# summaries go to gpurun_out/sanitize_*.log -> profiles/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -q -m gpu -x --tb=line -p no:cacheprovider"
run() {  # name tool timeout pytest-args...
  local name=$1 tool=$2 tmo=$3; shift 3
  echo "== $name ($tool)"
  timeout $tmo $SAN --tool $tool --launch-timeout 0 --error-exitcode 99 --print-limit 20 $PY "$@" > gpurun_out/sanitize_$name.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Race reported" gpurun_out/sanitize_$name.log | tail -6
}
run mem_pointwise memcheck 600 tests/test_kernels_gpu.py -k "bn_ or philox or keep_bit or reparam or kl or mse or intro_loss or relu_drop"
run mem_thin memcheck 600 tests/test_kernels_gpu.py -k "c1_to_cn or cn_to_c1 or wgrad_c1"
run mem_conv memcheck 1200 tests/test_conv_gpu.py -k "test_fprop or test_wgrad or test_upconv_fprop_bn_fused_stats or test_upconv_dgrad or test_conv_bn_fused_stats"
run race_pointwise racecheck 600 tests/test_kernels_gpu.py -k "bn_train_coeffs or bn_act_fwd_bwd or keep_bit"
run race_conv racecheck 1200 tests/test_conv_gpu.py -k "test_conv_bn_fused_stats or test_upconv_fprop_bn_fused_stats or test_upconv_wgrad"
