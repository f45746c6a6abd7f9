#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call18.txt
: > $out
echo "== train ops" >> $out
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q 2>&1 | tail -1 >> $out
for cfg in "2 1" "4 1" "4 -1"; do
  set -- $cfg
  echo "== attention backward: IMAGD_BWD_DQ_STAGES=$1 IMAGD_BWD_DKV2=$2" >> $out
  IMAGD_BWD_DQ_STAGES=$1 IMAGD_BWD_DKV2=$2 timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep -E "hd=40" | grep -v forward >> $out
  IMAGD_BWD_DQ_STAGES=$1 IMAGD_BWD_DKV2=$2 timeout 300 python -m pytest tests/test_train_ops_gpu.py -q -k "attention" 2>&1 | tail -1 >> $out
done
echo "== whole-step parity (defaults)" >> $out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -s -k "2-40-32 or graphed" 2>&1 | grep -E "grad rel|loss|passed|failed|Error" >> $out
echo "== train bench (graph), defaults" >> $out
timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_train_graph2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_train_graph2.json')); print({k:d['train'][k] for k in ('step_mode','samples_per_s','ms_per_step','e2e_samples_per_s','gpu_launches_per_step','model_frac_of_sustained_bf16')})" >> $out 2>&1
echo "== train bench (graph), DKV2=-1" >> $out
IMAGD_BWD_DKV2=-1 timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d['train'][k] for k in ('samples_per_s','ms_per_step')})" >> $out 2>&1
cat $out | cut -c1-200
