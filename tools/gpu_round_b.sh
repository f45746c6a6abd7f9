#!/bin/bash
# GN chunking sweep (env knobs) on the replayed step + correctness of the non-default chunking
mkdir -p gpurun_out
IMAGD_GN_PX=8 IMAGD_GN_MAXCHUNKS=128 timeout 200 python -m pytest tests/test_norm_elementwise_gpu.py -m gpu -q --timeout 150 2>&1 | tail -2
for cfg in "32 64" "16 64" "16 128" "8 128" "8 148"; do
  set -- $cfg
  for B in 1 8; do
    echo "GN_PX=$1 MAXCHUNKS=$2 B=$B: $(IMAGD_GN_PX=$1 IMAGD_GN_MAXCHUNKS=$2 B=$B timeout 200 python tools/step_timing.py 2>&1 | grep 'graph-replayed')"
  done
done
