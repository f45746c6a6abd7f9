#!/bin/bash
# First GPU validation of the training-step kernels: every test group in its own process (a trapped kernel poisons the
# CUDA context of its process only), output collected under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/train_call1.txt
: > $out
for k in "test_layout_kernels" "test_groupnorm_backward" "test_layernorm_backward" "test_activations_and_geglu" "test_mse_and_adamw" \
         "test_cross_attention_backward" "test_attention_backward_matches_autograd and 8-40-2-256" "test_attention_backward_matches_autograd and 8-40-2-200" \
         "test_attention_backward_matches_autograd and 8-80" "test_attention_backward_matches_autograd and 8-160-2" \
         "test_attention_backward_matches_autograd and 8-160-1" "test_attention_backward_matches_autograd and 8-40-2-384" \
         "test_attention_backward_matches_autograd and 12-64"; do
  echo "== $k" >> $out
  timeout 300 python -m pytest tests/test_train_ops_gpu.py -q -x -k "$k" 2>&1 | tail -25 >> $out
done
cat $out
