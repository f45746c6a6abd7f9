"""Where does a replayed denoising step go? Times (CUDA events): 50 graph replays, one eager step, the garment pass,
and the per-node floor of a CUDA graph made of tiny kernels. Run on the GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from imagdressing_b200 import ops

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "1"))
pipe = bench.build_product(dev)
x = bench.synth_inputs(B, dev)
for _ in range(2):
    bench.run_pipe(pipe, x)
eng = pipe._engine
st = next(iter(eng._states.values()))


def ev(fn, n=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


st["step_ptr"].zero_()
t_replay = ev(lambda: st["graph"].replay(), 50)
st["step_ptr"].zero_()
t0 = time.perf_counter()
t_eager = ev(lambda: eng._step(st), 3)
st["step_ptr"].zero_()
t_garment = ev(lambda: eng.garment_features(x["garment"], x["gtok"]), 3)
t_full = ev(lambda: bench.run_pipe(pipe, x), 2)
print(f"B={B}: graph-replayed step {t_replay:.3f} ms ({st['graph_launches']} kernels) | eager step {t_eager:.3f} ms | "
      f"garment pass {t_garment:.3f} ms | full image {t_full:.1f} ms")

# graph floor: N tiny dependent kernels
lat = torch.zeros(1, 4, 8, 8, device=dev)
eps = torch.zeros(1, 4, 8, 8, device=dev)
coef = torch.ones(4096, 4, device=dev)
sp = torch.zeros(2, dtype=torch.int32, device=dev)
ops.cfg_ddim_step(eps, None, 1.0, lat, coef, sp)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(400):
        ops.cfg_ddim_step(eps, None, 1.0, lat, coef, sp)
sp.zero_()
t_floor = ev(lambda: (sp.zero_(), g.replay()), 5)
print(f"graph of 400 tiny kernels: {t_floor:.3f} ms -> {t_floor / 400 * 1e3:.2f} us per node")
