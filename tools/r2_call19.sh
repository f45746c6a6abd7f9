#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call19.txt
: > $out
for d in 1 2; do
  echo "== attention backward: IMAGD_BWD_DQ_DSB=$d" >> $out
  IMAGD_BWD_DQ_DSB=$d timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep -E "hd=40" | grep -v forward >> $out
  IMAGD_BWD_DQ_DSB=$d timeout 300 python -m pytest tests/test_train_ops_gpu.py -q -k "attention" 2>&1 | tail -1 >> $out
done
echo "== train bench (graph), defaults" >> $out
timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d['train'][k] for k in ('samples_per_s','ms_per_step','model_frac_of_sustained_bf16')})" >> $out 2>&1
cat $out | cut -c1-200
