#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call9.txt
: > $out
for p in 0 1 2 3 4; do
  echo "== attention PP1 poly $p" >> $out
  IMAGD_ATTN_POLY=$p B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" >> $out
done
echo "== attention tests with poly 2" >> $out
IMAGD_ATTN_POLY=2 timeout 300 python -m pytest tests/test_attention_gpu.py -q 2>&1 | tail -2 >> $out
echo "== step A/B poly" >> $out
for p in 0 2 3; do
  for b in 1 8; do
    IMAGD_ATTN_POLY=$p B=$b timeout 300 python tools/step_timing.py 2>&1 | grep "graph-replayed" | sed "s/^/poly$p /" >> $out
  done
done
echo "== bench default" >> $out
timeout 1500 python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err
tail -c 3000 gpurun_out/bench_r02_default.json >> $out
cat $out | cut -c1-600
