#!/bin/bash
# Runs each GPU test file in its own process (a device-side trap in one file must not
# poison the others) under a timeout; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 420 python -m pytest "$f" -q -m gpu --no-header -p no:cacheprovider --timeout 120 > "gpurun_out/${name}.log" 2>&1
  r=$?
  echo "== $f exit $r"
  tail -n 25 "gpurun_out/${name}.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
