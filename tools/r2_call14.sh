#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call14.txt
: > $out
echo "== graphed step == eager step" >> $out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -s -k graphed 2>&1 | grep -E "losses|passed|failed|Error" >> $out
for pk in 0 1 2; do
  echo "== attention PP1 P-packing mode $pk (0 F2FP, 1 truncating byte permute, 2 round + permute)" >> $out
  IMAGD_ATTN_PACK=$pk B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" >> $out
  IMAGD_ATTN_PACK=$pk timeout 300 python -m pytest tests/test_attention_gpu.py -q 2>&1 | tail -1 >> $out
done
echo "== packing 1 + poly 2" >> $out
IMAGD_ATTN_PACK=1 IMAGD_ATTN_POLY=2 B=1,8 timeout 300 python tools/attn_bench.py 2>&1 | grep "hd=40" >> $out
for pk in 0 1; do
  for b in 1 8; do
    IMAGD_ATTN_PACK=$pk B=$b timeout 300 python tools/step_timing.py 2>&1 | grep "graph-replayed" | sed "s/^/pack$pk /" >> $out
  done
done
echo "== ncu: pp1 with packing 1" >> $out
IMAGD_ATTN_PACK=1 B=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pp1 -c 1 -o gpurun_out/r02_attn_pp1_pack1 -f python tools/attn_bench.py > gpurun_out/ncu_pack1.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_attn_pp1_pack1.ncu-rep >> $out 2>&1
cat $out | cut -c1-300
