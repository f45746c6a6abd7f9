#!/bin/bash
# ncu launch list (per-launch device time; cold-cache + serialised: compare SHARES) of 2 DDIM steps, eager launches.
mkdir -p gpurun_out
IMAGD_NO_GRAPH=1 IMAGD_DDIM_STEPS=2 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline \
  > gpurun_out/launches_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
