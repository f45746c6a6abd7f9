#!/bin/bash
# 2-GPU call: training-kernel re-validation (new GroupNorm / LayerNorm backward), then the data-parallel training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call10.txt
: > $out
echo "== norm backward kernels (rewritten)" >> $out
timeout 300 python -m pytest tests/test_train_ops_gpu.py -q -k "groupnorm or layernorm" 2>&1 | tail -3 >> $out
echo "== whole-step parity" >> $out
timeout 600 python -m pytest tests/test_train_step_gpu.py -q -s -k "2-16-16" 2>&1 | grep -E "grad rel|loss:|passed|failed|Error" >> $out
echo "== train bench N=1" >> $out
timeout 600 python bench.py --workload train --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-900 >> $out
echo "== train bench N=2 (NCCL gradient all-reduce from the bucket hooks)" >> $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 3 --warmup 3 2>&1 | grep -E "^\{|Error|error" | cut -c1-900 >> $out
cat $out
