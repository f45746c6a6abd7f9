"""One eager denoising step (CFG batch of B images) inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IMAGD_DDIM_STEPS"] = os.environ.get("IMAGD_DDIM_STEPS", "2")
import torch

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
torch.cuda.profiler.start()
eng._step(st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
