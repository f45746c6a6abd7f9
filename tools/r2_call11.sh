#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TRAIN=1 EPI=0 BATCHES=4 timeout 1500 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_train.log 2>&1
tail -5 gpurun_out/gemm_sweep_train.log
mv gpurun_out/gemm_tuning.inc gpurun_out/gemm_tuning_train.inc
mv gpurun_out/gemm_tuning.json gpurun_out/gemm_tuning_train.json
python - <<'PY'
import json
t=json.load(open("gpurun_out/gemm_tuning_train.json"))
tot_best=sum(v["us"] for v in t.values()); tot_auto=sum(v["auto_us"] for v in t.values())
print(len(t),"problems; sum best",round(tot_best),"us vs rule",round(tot_auto),"us (unweighted by launch count)")
PY
