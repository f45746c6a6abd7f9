#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_step_kernels_b1 -f python tools/kernel_ncu2.py > gpurun_out/ncu_step_kernels.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r02_step_kernels_b1.ncu-rep > gpurun_out/r02_ncu_step_kernels_b1.txt 2>&1
grep -E "Kernel Name|time_duration|tensor_cycles|pipe_xu|dram__bytes" gpurun_out/r02_ncu_step_kernels_b1.txt | cut -c1-160
