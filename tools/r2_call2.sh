#!/bin/bash
# Round-2 GPU call 2: P-in-TMEM attention (parity + timing + ncu), new defaults (LN fold / phase convs / bulk residual),
# the 50-step full-chain parity tests, step A/B.
mkdir -p gpurun_out
{
echo "== attention tests, P in TMEM (default)"
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_product_golden_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -3
echo "== attention tests, P in shared memory (IMAGD_ATTN_PTMEM=0)"
IMAGD_ATTN_PTMEM=0 timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -3
echo "== attention timing"
IMAGD_ATTN_PTMEM=0 timeout 120 python tools/attn_bench.py
IMAGD_ATTN_PTMEM=1 timeout 120 python tools/attn_bench.py
echo "== other GPU tests under the new defaults"
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_norm_elementwise_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_ln_fold_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -3
echo "== full-chain parity (50 steps, 512x512 and 768x576 inpaint)"
timeout 900 python -m pytest tests/test_full_chain_gpu.py -m gpu -x -q -s --timeout 800 2>&1 | tail -8
echo "== step A/B"
timeout 600 python tools/ab_step.py --batches=1,8 "r1:IMAGD_FOLD_LN=0,IMAGD_UPCONV_PHASE=0,IMAGD_GEMM_BULK_RES=0,IMAGD_ATTN_PTMEM=0" "new-ptmem0:IMAGD_ATTN_PTMEM=0" "new:" "new+persist:IMAGD_GEMM_PERSISTENT=1"
} 2>&1 | tee gpurun_out/call2.txt
echo "== ncu attention (P in TMEM)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 3 -c 1 -o gpurun_out/attn_v26 python tools/kernel_ncu.py > gpurun_out/ncu_attn26.log 2>&1
tail -2 gpurun_out/ncu_attn26.log
