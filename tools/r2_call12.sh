#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call12.txt
: > $out
echo "== train ops (vectorised transposes, weight layout kernel, bf16 fold outputs)" >> $out
timeout 600 python -m pytest tests/test_train_ops_gpu.py -q 2>&1 | tail -3 >> $out
echo "== whole-step parity + optimizer" >> $out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -s 2>&1 | grep -E "grad rel|loss|passed|failed|Error" >> $out
echo "== train bench N=1 (before the wgrad sweep is merged)" >> $out
timeout 600 python bench.py --workload train --steps 5 --warmup 5 2>&1 | tail -1 | cut -c1-420 >> $out
echo "== sweep of the training step's GEMM problems" >> $out
TRAIN=1 EPI=0 BATCHES=4 timeout 1500 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_train.log 2>&1
tail -3 gpurun_out/gemm_sweep_train.log >> $out
mv gpurun_out/gemm_tuning.inc gpurun_out/gemm_tuning_train.inc
mv gpurun_out/gemm_tuning.json gpurun_out/gemm_tuning_train.json
python - <<'PY' >> $out
import json
t=json.load(open("gpurun_out/gemm_tuning_train.json"))
tot_best=sum(v["us"] for v in t.values()); tot_auto=sum(v["auto_us"] for v in t.values())
print(len(t),"problems; sum best",round(tot_best),"us vs current choice",round(tot_auto),"us (unweighted by launch count)")
PY
cat $out | cut -c1-500
