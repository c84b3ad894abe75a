#!/bin/bash
# submit.sh <out-file> <timeout> <command...>: gpurun with retries while the pod answers "transient" (nothing charged)
OUT=$1; TMO=$2; shift 2
for i in 1 2 3 4 5 6; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $OUT 2>&1
  if grep -q "status=transient\|no box\|busy" $OUT && ! grep -q "status=ok" $OUT; then sleep 120; continue; fi
  break
done
tail -5 $OUT
