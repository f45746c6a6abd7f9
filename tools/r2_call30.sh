#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_call30_final_default.txt
: > $out
echo "== GroupNorm tests, shipped default (IMAGD_GN_CLUSTER unset)" >> $out
timeout 80 python -m pytest tests/test_norm_elementwise_gpu.py -q -k "groupnorm" 2>&1 | tail -2 >> $out
echo "== smoke(): 2-step inpaint pipeline + one training micro-step, both against the oracle" >> $out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4 >> $out
echo "== bench.py --no-extras --no-cpu-baseline" >> $out
timeout 120 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_b1_gn_cluster.json
python - >> $out <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_b1_gn_cluster.json"))
print({k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "gpu_launches")}, d.get("e2e"), d.get("roofline"))
PY
cat $out | cut -c1-400
