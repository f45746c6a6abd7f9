"""GroupNorm(+SiLU) launch time over the UNet's GroupNorm shapes, per kernel choice (IMAGD_GN_CLUSTER = 0 / 1 / 2; the library
reads the knob once, so each mode runs in its own process). CUDA events around 10 replays of a 20-launch CUDA graph per shape
after a warm-up; prints one table and the launch-weighted total for one denoising step.

    python tools/gn_bench.py [--batch 1]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (H, W, C, launches per step): ResnetBlock2D norm1 / norm2 and Transformer2DModel.norm of the SD1.5 UNet at 64 x 64 latents
SHAPES = [(64, 64, 320, 9), (32, 32, 320, 1), (32, 32, 640, 9), (16, 16, 640, 1), (16, 16, 1280, 10), (8, 8, 1280, 11),
          (8, 8, 2560, 2), (16, 16, 2560, 2), (16, 16, 1920, 1), (32, 32, 1920, 1), (32, 32, 1280, 1), (32, 32, 960, 1),
          (64, 64, 960, 1), (64, 64, 640, 2)]


def child(batch: int) -> None:
    import torch

    sys.path.insert(0, ROOT)
    from imagdressing_b200 import ops

    dev = torch.device("cuda:0")
    NB = 2 * batch
    total = 0.0
    for H, W, C, n in SHAPES:
        x = torch.randn(NB, H, W, C, device=dev).bfloat16()
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        out = torch.empty_like(x)
        for _ in range(5):
            ops.groupnorm(x, gamma, beta, 32, 1e-5, silu=True, out=out)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()  # 20 launches per replay: the python call (~11 us) would hide the kernel otherwise
        with torch.cuda.graph(graph):
            for _ in range(20):
                ops.groupnorm(x, gamma, beta, 32, 1e-5, silu=True, out=out)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1000.0 / 200
        total += us * n
        print(f"  [{NB},{H}x{W},{C}] x{n}: {us:7.2f} us", flush=True)
    print(f"  launch-weighted total per step: {total:8.1f} us", flush=True)


if __name__ == "__main__":
    batch = 1
    for i, a in enumerate(sys.argv):
        if a == "--batch":
            batch = int(sys.argv[i + 1])
    if os.environ.get("GN_BENCH_CHILD"):
        child(batch)
    else:
        for mode in ("0", "1", "2"):
            print(f"IMAGD_GN_CLUSTER={mode} (batch {batch})", flush=True)
            env = dict(os.environ, IMAGD_GN_CLUSTER=mode, GN_BENCH_CHILD="1")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--batch", str(batch)], env=env, capture_output=True,
                               text=True, timeout=300)
            print(r.stdout + (r.stderr[-600:] if r.returncode else ""), flush=True)
