#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call22.txt
: > $out
echo "== attention tests, ping-pong variant 3" >> $out
IMAGD_ATTN_PP_VARIANT=3 timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_product_golden_gpu.py -q -x 2>&1 | tail -5 >> $out
echo "== timing: variant 1 vs 3 (strict hand-over), 3 free-running" >> $out
IMAGD_ATTN_PP_VARIANT=1 B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" | sed "s/^/v1 /" >> $out
IMAGD_ATTN_PP_VARIANT=3 B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" | sed "s/^/v3 /" >> $out
IMAGD_ATTN_PP_VARIANT=3 IMAGD_ATTN_PP_SYNC=0 B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" | sed "s/^/v3-free /" >> $out
cat $out | cut -c1-200
