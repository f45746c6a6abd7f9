#!/bin/bash
# Round-2 GPU call 6: PDL modes on the step, the whole GPU suite under the final defaults, bench with edge sub-lines.
mkdir -p gpurun_out
{
echo "== step A/B: programmatic dependent launch"
timeout -s KILL 600 python tools/ab_step.py --batches=1,8 "pdl0:" "pdl1:IMAGD_PDL=1" "pdl2:IMAGD_PDL=2"
echo "== PDL=2 correctness (pipeline + unet tests)"
IMAGD_PDL=2 timeout -s KILL 600 python -m pytest tests/test_pipeline_gpu.py tests/test_gemm_gpu.py tests/test_attention_gpu.py -m gpu -x -q --timeout 500 2>&1 | tail -2
echo "== whole GPU suite"
( time timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -4 ) 2>&1
} 2>&1 | tee gpurun_out/call6.txt
echo "== bench"
timeout -s KILL 900 python bench.py > gpurun_out/bench_r02_b1.json 2> gpurun_out/bench_r02_b1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r02_b1.json; tail -3 gpurun_out/bench_r02_b1.err
