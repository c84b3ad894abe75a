#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in off 4 1; do
  if [ $v = off ]; then unset SIVAE_N256; else export SIVAE_N256=$v; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kt_$v.txt > gpurun_out/n256_$v.json 2> gpurun_out/n256_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/n256_$v.json')); print('N256=$v', round(d['ms_per_step'],2), round(d['value'],1), d['loss'])"
  grep "10, 12, 10, 256, 256" gpurun_out/kt_$v.txt
done
