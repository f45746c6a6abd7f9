#!/bin/bash
# First GPU call of round 2: validate the r2-prep items cheapest-first, then A/B them on the step. ~15 GPU-minutes:
# run it as  gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'  and read gpurun_out/ab_step.txt.
mkdir -p gpurun_out
set -o pipefail
echo "== default configuration"
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_norm_elementwise_gpu.py tests/test_attention_gpu.py \
  tests/test_product_golden_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -3
echo "== LayerNorm fold"
timeout 400 python -m pytest tests/test_ln_fold_gpu.py -m gpu -x -q --timeout 300 -s 2>&1 | tail -8
echo "== bulk store + bulk residual forced on"
IMAGD_GEMM_BULK_STORE=1 IMAGD_GEMM_BULK_RES=1 timeout 200 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -3
echo "== persistent GEMM"
IMAGD_GEMM_PERSISTENT=1 timeout 300 python -m pytest tests/test_gemm_persist_gpu.py tests/test_gemm_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -4
echo "== step A/B (batch 1: every knob; batch 8: the ones aimed at multi-wave grids)"
timeout 700 python tools/ab_step.py --batches=1 "base:IMAGD_GEMM_BULK_STORE=0" "rule:" "fold:IMAGD_FOLD_LN=1" \
  "upconv:IMAGD_UPCONV_PHASE=1" "fold+upconv:IMAGD_FOLD_LN=1,IMAGD_UPCONV_PHASE=1" "pdl:IMAGD_PDL=1" 2>&1 | tee gpurun_out/ab_step.txt
timeout 700 python tools/ab_step.py --batches=8 "base:IMAGD_GEMM_BULK_STORE=0" "rule:" "bulkres:IMAGD_GEMM_BULK_RES=1" \
  "persist:IMAGD_GEMM_PERSISTENT=1" "fold+upconv+bulkres:IMAGD_FOLD_LN=1,IMAGD_UPCONV_PHASE=1,IMAGD_GEMM_BULK_RES=1" 2>&1 | tee -a gpurun_out/ab_step.txt
# attention exp2 offload (compile-time knob): rebuild with 2 of every 8 exponentials on the FMA pipe, test, time, restore
echo "== attention exp2 polynomial offload (IMAGD_ATTN_POLY=2)"
touch imagdressing_b200/csrc/attention_tc.cu && make -s -C imagdressing_b200/csrc EXTRA=-DIMAGD_ATTN_POLY=2 > /dev/null
timeout 200 python -m pytest tests/test_attention_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -2
timeout 300 python tools/ab_step.py --batches=1,8 "poly2:" 2>&1 | tee -a gpurun_out/ab_step.txt
touch imagdressing_b200/csrc/attention_tc.cu && make -s -C imagdressing_b200/csrc > /dev/null
