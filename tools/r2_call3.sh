#!/bin/bash
# Round-2 GPU call 3: ping-pong attention kernel (parity, timing, ncu), GroupNorm shifted statistics, VAE support
# kernels, oracle anchor at full width, smoke(), bench.py with the new contract fields.
mkdir -p gpurun_out
{
echo "== attention tests, ping-pong kernel (default)"
timeout -s KILL 300 python -m pytest tests/test_attention_gpu.py tests/test_product_golden_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -4
echo "== attention timing"
IMAGD_ATTN_PP=0 timeout -s KILL 120 python tools/attn_bench.py 2>&1 | grep "hd=40"
IMAGD_ATTN_PP=1 timeout -s KILL 120 python tools/attn_bench.py 2>&1 | grep "hd=40"
echo "== norm / elementwise / VAE-support kernels, oracle anchor"
timeout -s KILL 600 python -m pytest tests/test_norm_elementwise_gpu.py tests/test_oracle_anchor_gpu.py -m gpu -x -q -s --timeout 500 2>&1 | grep -E "passed|failed|rel-L2|Error|error|taps" | tail -12
echo "== step A/B"
timeout -s KILL 400 python tools/ab_step.py --batches=1,8 "pp0:IMAGD_ATTN_PP=0" "pp1:"
echo "== smoke"
timeout -s KILL 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== unet + pipeline tests under the ping-pong kernel"
timeout -s KILL 900 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call3.txt
echo "== bench (new contract)"
timeout -s KILL 900 python bench.py > gpurun_out/bench_r02_b1.json 2> gpurun_out/bench_r02_b1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r02_b1.json; tail -3 gpurun_out/bench_r02_b1.err
echo "== ncu attention (ping-pong)"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/attn_pp python tools/kernel_ncu.py > gpurun_out/ncu_attn_pp.log 2>&1
tail -2 gpurun_out/ncu_attn_pp.log
