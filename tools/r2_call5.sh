#!/bin/bash
# Round-2 GPU call 5: ping-pong attention variant x sync matrix; CLIP encoders; causal attention; GEMM table re-sweep.
mkdir -p gpurun_out
{
for v in 1 2; do for s in 1 0; do
  echo "== attention PP variant $v sync $s"
  IMAGD_ATTN_PP_VARIANT=$v IMAGD_ATTN_PP_SYNC=$s timeout -s KILL 200 python -m pytest tests/test_attention_gpu.py -m gpu -x -q --timeout 150 2>&1 | tail -1
  IMAGD_ATTN_PP_VARIANT=$v IMAGD_ATTN_PP_SYNC=$s timeout -s KILL 120 python tools/attn_bench.py 2>&1 | grep "hd=40"
done; done
echo "== CLIP + causal attention"
timeout -s KILL 600 python -m pytest tests/test_clip_gpu.py -m gpu -x -q -s --timeout 500 2>&1 | grep -E "passed|failed|rel-L2|Error|error" | tail -8
} 2>&1 | tee gpurun_out/call5.txt
echo "== GEMM table re-sweep (batches 1, 8)"
BATCHES=1,8 timeout -s KILL 900 python tools/gemm_sweep.py > gpurun_out/sweep_r02.log 2>&1; tail -3 gpurun_out/sweep_r02.log
