#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call21.txt
: > $out
echo "== inference attention: softmax warps spinning on s_full (IMAGD_ATTN_SPIN)" >> $out
for sp in 0 1; do
  IMAGD_ATTN_SPIN=$sp B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" | sed "s/^/spin$sp /" >> $out
done
echo "== full GPU suite" >> $out
start=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=12 2>&1 | tail -25 >> $out
echo "wall=$(( $(date +%s) - start )) s" >> $out
cat $out | cut -c1-220
