"""Per-kernel device time of ONE eager denoising step via torch.profiler (CUPTI) — the cheap development profile
(the judged evidence is the ncu launch list in profiles/)."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "1"))
pipe = bench.build_product(dev)
x = bench.synth_inputs(B, dev)
bench.run_pipe(pipe, x)
eng = pipe._engine
st = next(iter(eng._states.values()))
st["step_ptr"].zero_()
eng._step(st)
torch.cuda.synchronize()
st["step_ptr"].zero_()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng._step(st)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
rows = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.device_time > 0:
        name = re.sub(r"\(.*", "", e.name).replace("imagd::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += e.device_time
        rows.append((e.device_time, name))
tot = sum(v[1] for v in agg.values())
print(f"B={B} one eager step: {len(rows)} kernels, {tot / 1e3:.3f} ms of kernel time")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% n={c:4d} avg={t / c:8.2f} us  {k[:100]}")
