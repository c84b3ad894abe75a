#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loss_curve.py -q -m gpu -x -s -k fc > gpurun_out/fc_curve.log 2>&1
echo "fc loss curve rc=$?"; grep -v "^step" gpurun_out/fc_curve.log | tail -14
timeout 200 python tools/ncu_linear.py && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"linear_(fwd|dgrad|wgrad)_kernel" -s 6 -c 6 \
  -o gpurun_out/r01f_linear -f python tools/ncu_linear.py > gpurun_out/ncu_linear.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_linear.log; ls -la gpurun_out/*.ncu-rep
