"""Tiny driver for `ncu --set full`: launches the level-0 hybrid attention (CFG batch 2: one conditional sample with
the garment stream + one unconditional) and the level-0 3x3 conv a few times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from imagdressing_b200 import ops

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "1"))
L, C, heads, hd, NB = 4096, 320, 8, 40, 2 * B
qkv = torch.randn(NB, L, 3 * C, device=dev).bfloat16()
kvr = torch.randn(B, L, 2 * C, device=dev).bfloat16()
flat = lambda t: t.as_strided((t.shape[0] * t.shape[1], t.shape[2]), (t.stride(1), 1), t.storage_offset())
s0 = ops.kv_stream(flat(qkv[..., C:2 * C]), flat(qkv[..., 2 * C:]), L)
s1 = ops.kv_stream(flat(kvr[..., :C]), flat(kvr[..., C:]), L, n_query_samples=B)
out = torch.empty(NB * L, C, device=dev, dtype=torch.bfloat16)
x = torch.randn(NB, 64, 64, C, device=dev).bfloat16()
w = (torch.randn(C, 9 * C, device=dev) * 0.02).bfloat16()
y = torch.empty(NB, 64, 64, C, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(flat(qkv[..., :C]), NB, L, heads, hd, s0, s1, out=out)
    ops.conv3x3(x, w, out=y)
torch.cuda.synchronize()
