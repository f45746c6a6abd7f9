#!/bin/bash
# 2 GPUs: data-parallel training step inside a CUDA graph (bucket all-reduces captured), then the driver-format default bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call15.txt
: > $out
echo "== train bench N=2, CUDA graph (NCCL all-reduces captured)" >> $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload train --steps 5 --warmup 5 2>&1 | grep -E "^\{|Error|error|Traceback" | cut -c1-1500 >> $out
echo "rc=$?" >> $out
echo "== default bench N=2 (driver command)" >> $out
start=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "rc=$? wall=$(( $(date +%s) - start )) s" >> $out
grep -E "^\{" gpurun_out/bench_n2.log > gpurun_out/bench_r02_n2.json
python - <<'PY' >> $out 2>&1
import json
d=json.load(open("gpurun_out/bench_r02_n2.json"))
print({k:d.get(k) for k in ("value","n_gpus","ms_per_step","e2e","gpu_launches")})
for b in d.get("batches",[]): print({k:b[k] for k in ("batch_per_gpu","global_batch","images_per_s")})
for b in d.get("other_configs",[]): print({k:v for k,v in b.items() if k not in ("clocks","what")})
PY
cat $out | cut -c1-700
