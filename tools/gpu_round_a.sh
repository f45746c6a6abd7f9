#!/bin/bash
# one GPU call: 768x576 UNet parity + ncu --set full of the remaining kernels + configs[2]/[3] bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --timeout 300 -rA 2>&1 | tail -40 > gpurun_out/test_unet_gpu.log
grep -E "passed|failed|error|rel-L2" gpurun_out/test_unet_gpu.log | tail -12
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/kernels2 \
  python tools/kernel_ncu2.py > gpurun_out/kernels2.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/kernels2.ncu-rep > gpurun_out/kernels2_summary.txt 2>&1
timeout 500 python bench.py --batch 32 --workload ipa_controlnet --steps 2 --warmup 3 > gpurun_out/bench_cfg2_b32_ipa_controlnet.json 2> gpurun_out/bench_cfg2.err
echo "cfg2 rc=$?"; cut -c1-700 gpurun_out/bench_cfg2_b32_ipa_controlnet.json
timeout 400 python bench.py --batch 8 --workload inpaint --height 768 --width 576 --steps 2 --warmup 3 > gpurun_out/bench_cfg3_b8_inpaint_768x576.json 2> gpurun_out/bench_cfg3.err
echo "cfg3 rc=$?"; cut -c1-700 gpurun_out/bench_cfg3_b8_inpaint_768x576.json
tail -5 gpurun_out/bench_cfg2.err gpurun_out/bench_cfg3.err
