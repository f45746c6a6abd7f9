#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call25.txt
: > $out
timeout 900 python -m pytest tests/test_train_step_gpu.py -q -s -k "trajectory" 2>&1 | grep -E "trajectory|passed|failed|Error" >> $out
timeout 300 python -m pytest tests/test_product_golden_gpu.py -q 2>&1 | tail -1 >> $out
cat $out | cut -c1-300
