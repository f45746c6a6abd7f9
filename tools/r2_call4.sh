#!/bin/bash
# Round-2 GPU call 4: ping-pong attention v2 (separate P buffer, early QK), VAE on the kernels, GroupNorm large-mean cases.
mkdir -p gpurun_out
{
echo "== attention tests, ping-pong v2"
timeout -s KILL 300 python -m pytest tests/test_attention_gpu.py tests/test_product_golden_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -4
echo "== attention timing"
IMAGD_ATTN_PP=1 timeout -s KILL 120 python tools/attn_bench.py 2>&1 | grep "hd=40"
echo "== VAE + norm tests"
timeout -s KILL 600 python -m pytest tests/test_vae_gpu.py tests/test_norm_elementwise_gpu.py -m gpu -x -q -s --timeout 500 2>&1 | grep -E "passed|failed|rel-L2|Error|error" | tail -14
echo "== step A/B"
timeout -s KILL 400 python tools/ab_step.py --batches=1,8 "pp0:IMAGD_ATTN_PP=0" "pp2:"
echo "== smoke"
timeout -s KILL 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} 2>&1 | tee gpurun_out/call4.txt
echo "== ncu attention (ping-pong v2)"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/attn_pp2 python tools/kernel_ncu.py > gpurun_out/ncu_attn_pp2.log 2>&1
tail -2 gpurun_out/ncu_attn_pp2.log
