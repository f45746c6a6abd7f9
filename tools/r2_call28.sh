#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_call28_gn_cluster.txt
: > $out
echo "== GroupNorm tests, IMAGD_GN_CLUSTER=2 (cluster kernel wherever a slice fits)" >> $out
IMAGD_GN_CLUSTER=2 timeout 240 python -m pytest tests/test_norm_elementwise_gpu.py tests/test_train_ops_gpu.py -q -k "groupnorm" 2>&1 | tail -4 >> $out
echo "== microbenchmark" >> $out
timeout 200 python tools/gn_bench.py --batch 1 >> $out 2>&1
timeout 200 python tools/gn_bench.py --batch 8 2>&1 | grep -E "IMAGD|total" >> $out
echo "== replayed step A/B" >> $out
timeout 400 python tools/ab_step.py "base:IMAGD_GN_CLUSTER=0" "cluster1:IMAGD_GN_CLUSTER=1" --batches=1,8 >> $out 2>&1
echo "== pipeline parity under IMAGD_GN_CLUSTER=1" >> $out
IMAGD_GN_CLUSTER=1 timeout 240 python -m pytest tests/test_pipeline_gpu.py -q -x -k "base_pipeline or inpainting" 2>&1 | tail -2 >> $out
cat $out | cut -c1-220
