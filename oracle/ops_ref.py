"""Per-op fp32 torch references for the kernel parity tests (device-agnostic; run in fp32, TF32 off).

Each mirrors one entry point of include/imagd_b200.h and cites the reference arithmetic it stands for.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def gemm_ref(a, w, bias=None, rowvec=None, rows_per_group=0, residual=None, act="none", alpha=1.0):
    """epilogue(a @ w^T): nn.Linear / 1x1 conv (adapter/attention_processor.py:568-615) + fused adds."""
    y = alpha * (a.float() @ w.float().t())
    if bias is not None:
        y = y + bias.float()
    if rowvec is not None:
        g = torch.arange(y.shape[0], device=y.device) // rows_per_group
        y = y + rowvec.float()[g]
    if act == "silu":
        y = F.silu(y)
    elif act == "gelu":
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


def geglu_pack(w, bias=None):
    """Interleave a [2*F, K] GEGLU projection (value rows then gate rows, diffusers-0.24 GEGLU.proj) into the
    kernel's packed order: per 128 packed rows, 64 value rows followed by their 64 gate rows."""
    F2, K = w.shape
    Fh = F2 // 2
    assert Fh % 64 == 0
    val, gate = w[:Fh].reshape(Fh // 64, 64, K), w[Fh:].reshape(Fh // 64, 64, K)
    wp = torch.cat([val, gate], dim=1).reshape(F2, K).contiguous()
    bp = None
    if bias is not None:
        bv, bg = bias[:Fh].reshape(Fh // 64, 64), bias[Fh:].reshape(Fh // 64, 64)
        bp = torch.cat([bv, bg], dim=1).reshape(F2).contiguous()
    return wp, bp


def geglu_ref(a, w, bias=None):
    """hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)  (diffusers-0.24 GEGLU.forward, erf GELU)."""
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    h, g = y.chunk(2, dim=-1)
    return h * F.gelu(g)


def conv3x3_pack(w_oihw):
    """[Cout, Cin, 3, 3] -> tap-major [Cout, 9*Cin]."""
    co, ci = w_oihw.shape[:2]
    return w_oihw.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous()


def conv3x3_ref(x_nhwc, w_oihw, bias=None, rowvec=None, residual=None, stride=1):
    """nn.Conv2d(k=3, pad=1) on NHWC data (ResnetBlock2D.conv1/2, Down/Upsample2D.conv; diffusers-0.24)."""
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w_oihw.float(), bias.float() if bias is not None else None,
                 stride=stride, padding=1)
    if rowvec is not None:
        y = y + rowvec.float()[:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    return y


def sdpa_ref(q, k, v, heads, scale=None):
    """F.scaled_dot_product_attention over [B, L, heads*hd] tensors (adapter/attention_processor.py:582-591)."""
    B, Lq, C = q.shape
    hd = C // heads
    qh = q.float().view(B, Lq, heads, hd).transpose(1, 2)
    kh = k.float().view(B, -1, heads, hd).transpose(1, 2)
    vh = v.float().view(B, -1, heads, hd).transpose(1, 2)
    scale = scale if scale is not None else 1.0 / math.sqrt(hd)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, C)


def hybrid_attention_ref(q, k0, v0, heads, k1=None, v1=None, w0=1.0, w1=1.0, n1=None, scale=None):
    """w0*SDPA(q,k0,v0) + w1*SDPA(q,k1,v1) for samples [0,n1) (adapter/attention_processor.py:589-612)."""
    out = w0 * sdpa_ref(q, k0, v0, heads, scale)
    if k1 is not None:
        n1 = q.shape[0] if n1 is None else n1
        if k1.shape[0] == 1 and n1 > 1:
            k1, v1 = k1.expand(n1, -1, -1), v1.expand(n1, -1, -1)
        out[:n1] = out[:n1] + w1 * sdpa_ref(q[:n1], k1[:n1], v1[:n1], heads, scale)
    return out


def groupnorm_ref(x_tok, gamma, beta, groups, eps, silu):
    """nn.GroupNorm on token-major [NB, HW, C] (+SiLU): ResnetBlock2D.norm1/2, Transformer2DModel.norm."""
    NB, C = x_tok.shape[0], x_tok.shape[-1]
    x = x_tok.float().reshape(NB, -1, C).transpose(1, 2)
    y = F.group_norm(x, groups, gamma.float() if gamma is not None else None,
                     beta.float() if beta is not None else None, eps)
    if silu:
        y = F.silu(y)
    return y.transpose(1, 2).reshape(x_tok.shape)


def layernorm_ref(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma.float() if gamma is not None else None,
                        beta.float() if beta is not None else None, eps)


def timestep_embedding_ref(t, dim):
    """diffusers-0.24 get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def ddim_step_ref(eps_c, eps_u, g, x, a_t, a_p):
    """CFG + DDIMScheduler.step(eta=0) (IMAGDressing_v1_pipeline.py:521-532)."""
    eps = eps_u + g * (eps_c - eps_u) if eps_u is not None else eps_c
    x0 = (x - math.sqrt(1 - a_t) * eps) / math.sqrt(a_t)
    return math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
