#!/bin/bash
# round-end confirmation: every GPU test file, the contract bench line (B=1), B=8, and the ncu launch list of one step
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_gemm_gpu.py tests/test_attention_gpu.py tests/test_norm_elementwise_gpu.py \
  tests/test_product_golden_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py
timeout 400 python bench.py > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "bench b1 rc=$?"; cut -c1-400 gpurun_out/bench_b1.json
timeout 300 python bench.py --batch 8 --steps 3 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench b8 rc=$?"; cut -c1-300 gpurun_out/bench_b8.json
B=1 bash tools/profile_launches.sh
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
