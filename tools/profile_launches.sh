#!/bin/bash
# ncu launch list (per-launch device time; cold-cache + serialised: compare SHARES) of ONE eager denoising step.
mkdir -p gpurun_out
B=${B:-1} timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches_b${B:-1}.csv python tools/one_step.py > gpurun_out/launches_run.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_b${B:-1}.csv
