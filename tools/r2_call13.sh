#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call13.txt
: > $out
echo "== whole-step parity (both sizes in ONE process: cache keys must survive model replacement) + optimizer + graph" >> $out
timeout 1200 python -m pytest tests/test_train_step_gpu.py -q -s 2>&1 | grep -E "grad rel|loss|passed|failed|Error|error" >> $out
echo "== train ops" >> $out
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q 2>&1 | tail -1 >> $out
echo "== train bench N=1: eager, then CUDA graph" >> $out
IMAGD_TRAIN_GRAPH=0 timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 | cut -c1-330 >> $out
timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_train_graph.json
cut -c1-330 gpurun_out/bench_train_graph.json >> $out
python -c "
import json; d=json.load(open('gpurun_out/bench_train_graph.json')); print({k:d['train'][k] for k in ('step_mode','samples_per_s','ms_per_step','e2e_samples_per_s','model_frac_of_sustained_bf16','loss_first_steps')})" >> $out 2>&1
cat $out | cut -c1-400
