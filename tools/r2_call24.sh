#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call24.txt
: > $out
echo "== full GPU suite" >> $out
start=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 >> $out
echo "wall=$(( $(date +%s) - start )) s" >> $out
echo "== smoke" >> $out
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $out
echo "== default bench (driver command, N=1)" >> $out
start=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
echo "rc=$? wall=$(( $(date +%s) - start )) s" >> $out
python - <<'PY' >> $out 2>&1
import json
d=json.load(open("gpurun_out/bench_r02_final.json"))
print({k:d.get(k) for k in ("value","ms_per_step","e2e","gpu_launches","clocks","model_frac_of_sustained_bf16")})
print("roofline", {k:d["roofline"].get(k) for k in ("kernel","achieved","frac","traffic","step_share")})
for b in d.get("batches",[]): print({k:b[k] for k in ("batch_per_gpu","images_per_s","model_frac_of_sustained_bf16")})
for b in d.get("other_configs",[]): print({k:v for k,v in b.items() if k not in ("clocks","what")})
print(d.get("cpu_baseline"))
PY
echo "== reference arm" >> $out
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-400 >> $out
cat $out | cut -c1-330
echo "== reference arm, training step on the host cores" >> gpurun_out/call24.txt
timeout 900 python bench.py --impl reference --workload train --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-600 >> gpurun_out/call24.txt
tail -2 gpurun_out/call24.txt | cut -c1-600
