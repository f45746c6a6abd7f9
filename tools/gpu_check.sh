#!/bin/bash
# Run each GPU test file in its own process (a device trap poisons only that process); logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 180 -rA 2>&1 | tail -120 > "gpurun_out/$name.log"
  code=${PIPESTATUS[0]}
  echo "== $f -> exit $code"
  grep -E "passed|failed|error" "gpurun_out/$name.log" | tail -3
  [ $code -ne 0 ] && rc=1
done
exit $rc
