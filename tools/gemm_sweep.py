"""Tune the tcgen05 GEMM/conv variant choice: record every distinct GEMM problem the garment pass, the CFG-batched
denoising step and the ControlNet issue at the given per-GPU batch sizes, time every (N tile, ring depth, split-K)
variant on each (L2 flushed between launches, CUDA events, median of 5) and write the winners to
gpurun_out/gemm_tuning.json / .inc (copy the .inc to imagdressing_b200/csrc/gemm_tuning.inc and rebuild).

    BATCHES=1,2,8 python tools/gemm_sweep.py          # EPI=1 (default): realistic epilogues — bias + residual on the
                                                       # linears, bias + time-embedding row vector on the convs, bias on
                                                       # GEGLU; EPI=0: bare products (how the round-1 table was swept)
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IMAGD_NO_GRAPH"] = "1"
os.environ["IMAGD_DDIM_STEPS"] = "1"
import torch

import bench
from imagdressing_b200 import _lib, modeling, ops

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.int32)
BATCHES = [int(v) for v in os.environ.get("BATCHES", "1").split(",")]
EPI = os.environ.get("EPI", "1") == "1"
HW = int(os.environ.get("HW", "64"))
bench.HW = HW

# ---- 1. record the problems
TRAIN = os.environ.get("TRAIN", "0") == "1"  # TRAIN=1 EPI=0: the problems of one training step (forward, dgrad, wgrad)
if TRAIN:
    from imagdressing_b200 import train

    sd, opt, sched = bench.build_train(dev)
    lib.imagd_gemm_debug_log(1, None, 0)
    for B in BATCHES:
        xb = bench.synth_train_batch(B, dev, 0, False, int(os.environ.get("TH", "640")) // 8, int(os.environ.get("TW", "512")) // 8)
        train.train_step(sd, sched, optimizer=opt, **xb)
    torch.cuda.synchronize()
    del sd, opt
    pipe = controlnet = None
else:
    pipe = bench.build_product(dev)
    cn = modeling.UNet2DConditionModel  # noqa
    controlnet = modeling.ControlNetModel().to(dev, torch.bfloat16)
    modeling.init_synthetic_fast_(controlnet, 2)
    lib.imagd_gemm_debug_log(1, None, 0)
    for B in BATCHES:
        x = bench.synth_inputs(B, dev)
        bench.run_pipe(pipe, x)
        lat2 = torch.randn(2 * B, 4, HW, HW, device=dev)
        controlnet(lat2, torch.tensor([500.0], device=dev), torch.randn(2 * B, 77, 768, device=dev),
                   torch.rand(B, 3, HW * 8, HW * 8, device=dev), return_dict=False)
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 20)
n = lib.imagd_gemm_debug_log(0, buf, len(buf))
keys = [tuple(int(v) for v in ln.split()) for ln in buf.value.decode().strip().split("\n")]
print(f"{n} distinct GEMM problems for batches {BATCHES}")
del pipe, controlnet
torch.cuda.empty_cache()


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


SHALLOW = {64: 4, 128: 3, 160: 3, 256: 2}
DEEP = {64: 8, 128: 6, 160: 5, 256: 4}
entries, table = [], {}
for (taps, NB, H, W, Cin, N, geglu, m_tiles, kb_total, out_fp32) in sorted(set(keys)):
    if taps == 9:
        x = torch.randn(NB, H, W, Cin, device=dev).bfloat16()
        w = (torch.randn(N, 9 * Cin, device=dev) * 0.02).bfloat16()
        y = torch.empty(NB, H, W, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev) if EPI else None
        temb = torch.randn(NB, N, device=dev) if EPI else None
        fn = lambda: ops.conv3x3(x, w, out=y, bias=bias, rowvec=temb)
        flops = 2.0 * NB * H * W * 9 * Cin * N
    else:
        x = torch.randn(W, Cin, device=dev).bfloat16()
        w = (torch.randn(N, Cin, device=dev) * 0.02).bfloat16()
        y = torch.empty(W, N // 2 if geglu else N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
        act = ops.ACT_GEGLU if geglu else ops.ACT_NONE
        bias = torch.randn(N, device=dev) if EPI else None
        res = torch.randn(W, N, device=dev).bfloat16() if (EPI and not geglu and not out_fp32) else None
        fn = lambda: ops.gemm(x, w, out=y, act=act, out_fp32=bool(out_fp32), bias=bias, residual=res)
        flops = 2.0 * W * Cin * N
    row = {}
    for bn in ((128,) if geglu else (64, 128, 160, 256)):
        if bn > 64 and bn >= 2 * N:
            continue
        for stages in (SHALLOW[bn], DEEP[bn]):
            for S in (1, 2, 3, 4, 6, 8):
                if S > 1 and (geglu or kb_total // S < 8):
                    continue
                lib.imagd_gemm_debug_force(bn, stages, S)
                try:
                    row[f"{bn}/{stages}/{S}"] = round(timeit(fn), 1)
                except Exception:
                    row[f"{bn}/{stages}/{S}"] = None
    lib.imagd_gemm_debug_force(0, 0, 0)
    auto = round(timeit(fn), 1)
    best_t, best_k = min((v, k) for k, v in row.items() if v is not None)
    bn, stages, S = (int(v) for v in best_k.split("/"))
    entries.append((m_tiles, N, kb_total, geglu, bn, stages, S))
    table[f"{taps} {NB} {H} {W} {Cin} {N} {geglu}"] = dict(best=best_k, us=best_t, auto_us=auto, tflops=round(flops / best_t / 1e6, 1), all=row)
    print(f"taps={taps} NB={NB:3d} {H:3d}x{W:<5d} K={Cin * taps:6d} N={N:5d} g={geglu} m_tiles={m_tiles:4d}: best {best_k:>9s} "
          f"{best_t:7.1f} us {flops / best_t / 1e6:7.1f} TF/s | rule {auto:7.1f} us")

os.makedirs("gpurun_out", exist_ok=True)
json.dump(table, open("gpurun_out/gemm_tuning.json", "w"), indent=0)
with open("gpurun_out/gemm_tuning.inc", "w") as f:
    f.write("// {m_tiles, N, kb_total, geglu, N_tile, stages, splits} — generated by tools/gemm_sweep.py on a B200; "
            f"batches {BATCHES}, latent {HW}x{HW}\n")
    seen = set()
    for e in entries:
        if e[:4] in seen:
            continue
        seen.add(e[:4])
        f.write("{%d, %d, %d, %d, %d, %d, %d},\n" % e)
print("wrote gpurun_out/gemm_tuning.inc with", len(seen), "entries")
