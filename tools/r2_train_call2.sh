#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/train_call2.txt
: > $out
for k in "test_training_step_gradients_match_oracle and 2-16-16" "test_training_step_gradients_match_oracle and 2-40-32" "test_optimizer_step"; do
  echo "== $k" >> $out
  timeout 600 python -m pytest tests/test_train_step_gpu.py -q -x -s -k "$k" 2>&1 | grep -v Warning | tail -40 >> $out
done
cat $out
