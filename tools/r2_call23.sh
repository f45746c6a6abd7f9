#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call23.txt
: > $out
echo "== train ops (training forward through the ping-pong kernel; vectorised AdamW)" >> $out
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q 2>&1 | tail -1 >> $out
echo "== inference attention tests (pp1 epilogue touched)" >> $out
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_product_golden_gpu.py -q 2>&1 | tail -1 >> $out
echo "== training forward attention: ping-pong vs two-CTA kernel" >> $out
timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep "forward" | sed "s/^/pp /" >> $out
IMAGD_ATTN_TRAIN_PP=0 timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep "forward" | sed "s/^/2cta /" >> $out
echo "== whole-step parity + graphed" >> $out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -s -k "2-40-32 or graphed" 2>&1 | grep -E "grad rel|loss|passed|failed|Error" >> $out
echo "== train bench (graph)" >> $out
timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d['train'][k] for k in ('samples_per_s','ms_per_step','model_frac_of_sustained_bf16')})" >> $out 2>&1
cat $out | cut -c1-200
