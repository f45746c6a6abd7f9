#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call20.txt
: > $out
for cfg in "0 1" "1 1" "1 3" "0 3"; do
  set -- $cfg
  echo "== attention backward: IMAGD_BWD_SPIN=$1 IMAGD_BWD_DKV2=$2" >> $out
  IMAGD_BWD_SPIN=$1 IMAGD_BWD_DKV2=$2 timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep -E "hd=40" | grep -v forward >> $out
  IMAGD_BWD_SPIN=$1 IMAGD_BWD_DKV2=$2 timeout 300 python -m pytest tests/test_train_ops_gpu.py -q -k "attention" 2>&1 | tail -1 >> $out
done
echo "== other head dims with spin" >> $out
IMAGD_BWD_SPIN=1 timeout 300 python tools/attn_bwd_bench.py 2>&1 | grep -vE "hd=40" | grep -v forward >> $out
cat $out | cut -c1-200
