#!/bin/bash
# Round-2 GPU call 1: environment probe, pipe microbenchmarks, r2-prep validation + A/B.
mkdir -p gpurun_out
{
echo "== probe"; python - <<'PY'
import importlib
for m in ["diffusers", "accelerate", "torchvision", "safetensors", "PIL", "insightface", "onnxruntime", "xformers"]:
    try:
        mod = importlib.import_module(m); print(m, "OK", getattr(mod, "__version__", "?"))
    except Exception as e:
        print(m, "ABSENT", type(e).__name__, str(e)[:80])
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; free -g | head -2
echo "== ubench"; timeout 120 tools/ubench_pipes
} 2>&1 | tee gpurun_out/call1_probe.txt
bash tools/r2_first_call.sh 2>&1 | tee gpurun_out/call1_r2prep.txt
