"""Driver for `ncu --set full --profile-from-start off`: ONE profiled launch each of the remaining step kernels at
their batch-B level-0 / level-1 / level-2 shapes (CFG batch 2B): fused GroupNorm+SiLU (HBM-bound), LayerNorm,
the GEGLU GEMM and the FF-out GEMM (tcgen05), head-dim-80 and head-dim-160 attention. Warm-up launches are outside
the profiled range."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from imagdressing_b200 import ops

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "1"))
NB = 2 * B
r = lambda *s: torch.randn(*s, device=dev).bfloat16()
flat = lambda t: t.as_strided((t.shape[0] * t.shape[1], t.shape[2]), (t.stride(1), 1), t.storage_offset())

x0 = r(NB, 64, 64, 320)
g0, b0 = torch.randn(320, device=dev), torch.randn(320, device=dev)
x2 = r(NB, 16, 16, 1280)
g2, b2 = torch.randn(1280, device=dev), torch.randn(1280, device=dev)
tok = r(NB * 4096, 320)
w_geglu, bias_geglu = (torch.randn(2560, 320, device=dev) * 0.05).bfloat16(), torch.randn(2560, device=dev)
h_ff = r(NB * 4096, 1280)
w_out, bias_out = (torch.randn(320, 1280, device=dev) * 0.03).bfloat16(), torch.randn(320, device=dev)
res = r(NB * 4096, 320)


def attn_case(L, C, heads):
    qkv = r(NB, L, 3 * C)
    kvr = r(B, L, 2 * C)
    s0 = ops.kv_stream(flat(qkv[..., C:2 * C]), flat(qkv[..., 2 * C:]), L)
    s1 = ops.kv_stream(flat(kvr[..., :C]), flat(kvr[..., C:]), L, n_query_samples=B)
    out = torch.empty(NB * L, C, device=dev, dtype=torch.bfloat16)
    return lambda: ops.attention(flat(qkv[..., :C]), NB, L, heads, C // heads, s0, s1, out=out)


cases = [
    lambda: ops.groupnorm(x0, g0, b0, 32, 1e-5, silu=True),
    lambda: ops.groupnorm(x2, g2, b2, 32, 1e-5, silu=True),
    lambda: ops.layernorm(tok, g0, b0),
    lambda: ops.gemm(tok, w_geglu, bias=bias_geglu, act=ops.ACT_GEGLU),
    lambda: ops.gemm(h_ff, w_out, bias=bias_out, residual=res),
    attn_case(1024, 640, 8),
    attn_case(256, 1280, 8),
]
for _ in range(3):
    for c in cases:
        c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for c in cases:
    c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
