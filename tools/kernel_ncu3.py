"""Driver for `ncu --set full --profile-from-start off`: ONE profiled launch each of the training-step kernels at their
level-0 shapes of BASELINE.json configs[4] (micro-batch 4, 640x512 -> 5120 tokens x 320 channels per sample): training-mode
two-stream attention forward, its backward (dQ kernel, dK/dV kernel per stream), the conv weight-gradient GEMM with its
operand builders (transpose, transposed im2col), the conv data-gradient conv, GroupNorm / LayerNorm backward.
Warm-up launches are outside the profiled range."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from imagdressing_b200 import ops

dev = torch.device("cuda:0")
B, L, C, heads = 4, 5120, 320, 8
r = lambda *s: torch.randn(*s, device=dev).bfloat16()
qkv, kv1, d_out = r(B, L, 3 * C), r(B, L, 2 * C), r(B * L, C)
flat = qkv.view(B * L, 3 * C)
f1 = kv1.view(B * L, 2 * C)
s0 = ops.kv_stream(flat[:, C:2 * C], flat[:, 2 * C:], L)
s1 = ops.kv_stream(f1[:, :C], f1[:, C:], L, out_scale=1.0)
out, saved = ops.attention_train(flat[:, :C], B, L, heads, C // heads, s0, s1)
dqkv, dkv1 = torch.empty_like(flat), torch.empty_like(f1)

x = r(B, 80, 64, C)
dy = r(B, 80, 64, C)
wp = (torch.randn(C, 9 * C, device=dev) * 0.02).bfloat16()
gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
stats = torch.empty(B, 32, 2, device=dev)
ops.groupnorm(x, gamma, beta, 32, 1e-5, silu=True, stats_out=stats)
tok, dtok = r(B * L, C), r(B * L, C)


def wgrad():
    return ops.gemm(ops.transpose(dy.view(-1, C)), ops.im2col3x3_t(x))


cases = [
    lambda: ops.attention_train(flat[:, :C], B, L, heads, C // heads, s0, s1),
    lambda: ops.attention_bwd(flat[:, :C], d_out, B, L, heads, C // heads, s0, s1, saved, dq=dqkv[:, :C],
                              dkv0=(dqkv[:, C:2 * C], dqkv[:, 2 * C:]), dkv1=(dkv1[:, :C], dkv1[:, C:])),
    wgrad,
    lambda: ops.conv3x3(dy, wp),
    lambda: ops.groupnorm_bwd(x, dy, gamma, beta, 32, stats, True, True),
    lambda: ops.layernorm_bwd(tok, dtok, gamma, 1e-5, True),
]
for _ in range(2):
    for c in cases:
        c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for c in cases:
    c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
