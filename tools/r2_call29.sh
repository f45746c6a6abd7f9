#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02_call29_gn_cluster_bulk.txt
: > $out
echo "== GroupNorm tests, IMAGD_GN_CLUSTER=2 IMAGD_GN_BULK=1" >> $out
IMAGD_GN_CLUSTER=2 timeout 240 python -m pytest tests/test_norm_elementwise_gpu.py tests/test_train_ops_gpu.py -q -k "groupnorm" 2>&1 | tail -3 >> $out
echo "== microbenchmark (graph-replayed launches)" >> $out
timeout 240 python tools/gn_bench.py --batch 1 >> $out 2>&1
echo "== replayed step A/B" >> $out
timeout 400 python tools/ab_step.py "base:IMAGD_GN_CLUSTER=0" "cl1:IMAGD_GN_CLUSTER=1,IMAGD_GN_BULK=0" "cl1bulk:IMAGD_GN_CLUSTER=1,IMAGD_GN_BULK=1" "cl2bulk:IMAGD_GN_CLUSTER=2,IMAGD_GN_BULK=1" --batches=1 >> $out 2>&1
timeout 200 python tools/ab_step.py "base:IMAGD_GN_CLUSTER=0" "cl2bulk:IMAGD_GN_CLUSTER=2,IMAGD_GN_BULK=1" --batches=8 >> $out 2>&1
cat $out | cut -c1-220
