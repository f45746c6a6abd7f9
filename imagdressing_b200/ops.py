"""Tensor-level wrappers over the C ABI: torch owns device memory and the stream, the kernels do the arithmetic.

Every function takes CUDA tensors, passes raw pointers / strides / the current stream to libimagd_b200.so and
returns the output tensor. Nothing here computes on the CPU and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, Epilogue, KVStream

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a token-major view whose last dim is contiguous."""
    assert t.stride(-1) == 1, "last dim must be contiguous"
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    assert t.is_contiguous() or t.dim() == 2, "higher-rank activations must be contiguous"
    rows = 1
    for s in t.shape[:-1]:
        rows *= s
    return rows, t.shape[-1], t.shape[-1]


class LnFold:
    """LayerNorm folded into the consuming GEMM (include/imagd_b200.h): `stats` [M, ld, 2] fp32
    holds the producer's per-row {sum, sum of squares} partials (`parts` of them), `colsum` [N] the column sums of the
    gamma-scaled weight; the GEMM's `bias` must be the folded bias b + W beta."""

    __slots__ = ("stats", "parts", "dim", "eps", "colsum")

    def __init__(self, stats: torch.Tensor, parts: int, dim: int, eps: float, colsum: torch.Tensor):
        self.stats, self.parts, self.dim, self.eps, self.colsum = stats, int(parts), int(dim), float(eps), colsum


def gemm_tile_count_n(M: int, N: int, K: int) -> int:
    """How many row-statistics partials a producer GEMM of this shape writes per row."""
    n = _lib.load().imagd_gemm_tile_count_n(int(M), int(N), int(K))
    if n <= 0:
        _lib.check(n if n < 0 else -1, "imagd_gemm_tile_count_n")
    return n


def _epilogue(bias, rowvec, rows_per_group, residual, act, alpha, out_fp32) -> Epilogue:
    ep = Epilogue()
    ep.bias = _ptr(bias)
    ep.rowvec = _ptr(rowvec)
    ep.rowvec_ld = rowvec.stride(0) if rowvec is not None else 0
    ep.rows_per_group = int(rows_per_group)
    ep.act = int(act)
    ep.residual = _ptr(residual)
    ep.ldr = _rows2d(residual)[2] if residual is not None else 0
    ep.alpha = float(alpha)
    ep.out_fp32 = 1 if out_fp32 else 0
    return ep


def gemm(a: torch.Tensor, w: torch.Tensor, *, out: Optional[torch.Tensor] = None, bias=None, rowvec=None,
         rows_per_group: int = 0, residual=None, act: int = ACT_NONE, alpha: float = 1.0,
         out_fp32: bool = False, stats_out: Optional[torch.Tensor] = None, ln: Optional[LnFold] = None) -> torch.Tensor:
    """out[M, N] = epilogue(a[M, K] @ w[N, K]^T); a / w / residual bf16, bias / rowvec fp32.
    stats_out [M, ld, 2] fp32: also emit per-row {sum, sum of squares} of the rounded outputs, one slot per N tile.
    ln: the rows of `a` are the RAW input of a LayerNorm that has been folded into `w` / `bias` (LnFold)."""
    lib = _lib.load()
    M, K, lda = _rows2d(a)
    N, Kw = w.shape
    assert Kw == K and a.dtype == BF16 and w.dtype == BF16
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty(*a.shape[:-1], n_out, device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    ldd = _rows2d(out)[2]
    ep = _epilogue(bias, rowvec, rows_per_group, residual, act, alpha, out_fp32)
    if stats_out is not None:
        assert stats_out.dtype == torch.float32 and stats_out.dim() == 3 and stats_out.shape[0] == M and \
            stats_out.shape[2] == 2 and stats_out.is_contiguous()
        ep.row_stats_out = stats_out.data_ptr()
        ep.stats_ld = stats_out.shape[1]
    if ln is not None:
        assert ln.stats.dtype == torch.float32 and ln.stats.shape[0] == M and ln.stats.is_contiguous() and ln.dim == K
        assert ln.colsum.dtype == torch.float32 and ln.colsum.numel() == N and bias is not None
        ep.row_stats_in = ln.stats.data_ptr()
        ep.stats_in_ld = ln.stats.shape[1]
        ep.stats_parts = ln.parts
        ep.ln_dim = ln.dim
        ep.ln_eps = ln.eps
        ep.colsum = ln.colsum.data_ptr()
    rc = lib.imagd_gemm_bf16(a.data_ptr(), lda, w.data_ptr(), w.stride(0), out.data_ptr(), ldd, M, N, K,
                             ctypes.byref(ep), _stream())
    _lib.check(rc, "imagd_gemm_bf16")
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, out: Optional[torch.Tensor] = None, bias=None, rowvec=None,
            residual=None, act: int = ACT_NONE) -> torch.Tensor:
    """x: [NB, H, W, Cin] bf16 (token-major), w: [Cout, 9*Cin] tap-major. Stride 1, zero pad 1."""
    lib = _lib.load()
    NB, H, W, Cin = x.shape
    assert x.is_contiguous() and x.dtype == BF16
    Cout = w.shape[0]
    assert w.shape[1] == 9 * Cin
    if out is None:
        out = torch.empty(NB, H, W, Cout, device=x.device, dtype=BF16)
    ep = _epilogue(bias, rowvec, H * W, residual, act, 1.0, False)
    rc = lib.imagd_conv3x3_bf16(x.data_ptr(), Cin, NB, H, W, Cin, w.data_ptr(), out.data_ptr(), out.shape[-1], Cout,
                                ctypes.byref(ep), _stream())
    _lib.check(rc, "imagd_conv3x3_bf16")
    return out


def upconv3x3(x: torch.Tensor, w_phase: torch.Tensor, *, bias=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Upsample2D (nearest 2x) + 3x3 conv as four 2x2 phase convs on the low-resolution input (modeling.pack_upconv3x3).
    x: [NB, H, W, Cin] bf16, w_phase: [4*Cout, 4*Cin] bf16 -> [NB, 2H, 2W, Cout]."""
    lib = _lib.load()
    NB, H, W, Cin = x.shape
    assert x.is_contiguous() and x.dtype == BF16 and w_phase.dtype == BF16 and w_phase.shape[1] == 4 * Cin
    Cout = w_phase.shape[0] // 4
    if out is None:
        out = torch.empty(NB, 2 * H, 2 * W, Cout, device=x.device, dtype=BF16)
    ep = _epilogue(bias, None, 0, None, ACT_NONE, 1.0, False)
    rc = lib.imagd_upconv3x3_bf16(x.data_ptr(), Cin, NB, H, W, Cin, w_phase.data_ptr(), out.data_ptr(), out.shape[-1], Cout,
                                  ctypes.byref(ep), _stream())
    _lib.check(rc, "imagd_upconv3x3_bf16")
    return out


def kv_stream(k: torch.Tensor, v: torch.Tensor, length: int, *, sample_rows: int = 0, broadcast: bool = False,
              n_query_samples: int = 1 << 30, out_scale: float = 1.0) -> KVStream:
    """k / v: 2-D views [rows, C] whose row 0 is the first visited key of sample 0; `sample_rows` = rows between
    samples when that differs from `length` (a window of a longer context)."""
    s = KVStream()
    s.sample_rows = int(sample_rows)
    assert k.stride(-1) == 1 and v.stride(-1) == 1 and k.dtype == BF16 and v.dtype == BF16
    ld = k.stride(-2)
    assert v.stride(-2) == ld
    s.k, s.v, s.ld, s.len = k.data_ptr(), v.data_ptr(), ld, int(length)
    s.broadcast = 1 if broadcast else 0
    s.n_query_samples = int(min(n_query_samples, 1 << 30))
    s.out_scale = float(out_scale)
    return s


def attention(q: torch.Tensor, B: int, Lq: int, heads: int, head_dim: int, s0: KVStream,
              s1: Optional[KVStream] = None, *, sm_scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, causal: bool = False) -> torch.Tensor:
    """q: [B*Lq, >= heads*head_dim] bf16 view (row stride arbitrary). Returns [B*Lq, heads*head_dim] bf16.
    causal: query i attends to keys 0..i of stream 0 (CLIP text encoder); no second stream then."""
    lib = _lib.load()
    assert q.dtype == BF16 and q.stride(-1) == 1
    q_ld = q.stride(-2)
    if out is None:
        out = torch.empty(B * Lq, heads * head_dim, device=q.device, dtype=BF16)
    if sm_scale is None:
        sm_scale = head_dim ** -0.5
    if causal:
        assert s1 is None
        rc = lib.imagd_attention_causal_bf16(q.data_ptr(), q_ld, out.data_ptr(), out.stride(-2), B, Lq, heads, head_dim,
                                             ctypes.byref(s0), float(sm_scale), _stream())
        _lib.check(rc, "imagd_attention_causal_bf16")
        return out
    rc = lib.imagd_attention_bf16(q.data_ptr(), q_ld, out.data_ptr(), out.stride(-2), B, Lq, heads, head_dim,
                                  ctypes.byref(s0), ctypes.byref(s1) if s1 is not None else None, float(sm_scale),
                                  _stream())
    _lib.check(rc, "imagd_attention_bf16")
    return out


_gn_ws = {}


def _gn_workspace(device, nbytes: int) -> torch.Tensor:
    ws = _gn_ws.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)  # counters must start at 0
        _gn_ws[device] = ws
    return ws


def groupnorm(x: torch.Tensor, gamma, beta, groups: int, eps: float, *, silu: bool, out: Optional[torch.Tensor] = None,
              ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [NB, HW..., C] bf16 token-major (contiguous). GroupNorm over (HW, C/groups) per sample, optional SiLU."""
    lib = _lib.load()
    assert x.is_contiguous() and x.dtype == BF16
    NB, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (NB * C)
    if out is None:
        out = torch.empty_like(x)
    if ws is None:
        ws = _gn_workspace(x.device, lib.imagd_groupnorm_ws_bytes(NB, HW, C, groups))
    rc = lib.imagd_groupnorm_bf16(x.data_ptr(), C, out.data_ptr(), C, NB, HW, C, groups, _ptr(gamma), _ptr(beta),
                                  float(eps), 1 if silu else 0, ws.data_ptr(), _stream())
    _lib.check(rc, "imagd_groupnorm_bf16")
    return out


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    rows, C, ldx = _rows2d(x)
    assert x.dtype == BF16
    if out is None:
        out = torch.empty(*x.shape, device=x.device, dtype=BF16)
    rc = lib.imagd_layernorm_bf16(x.data_ptr(), ldx, out.data_ptr(), _rows2d(out)[2], rows, C, _ptr(gamma), _ptr(beta),
                                  float(eps), _stream())
    _lib.check(rc, "imagd_layernorm_bf16")
    return out


def concat_add(a: torch.Tensor, b: Optional[torch.Tensor] = None, *, res_a=None, res_b=None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cat([a (+res_a), b (+res_b)], -1) over token-major tensors of equal row count."""
    lib = _lib.load()
    rows, Ca, lda = _rows2d(a)
    Cb, ldb = 0, 0
    if b is not None:
        rb, Cb, ldb = _rows2d(b)
        assert rb == rows
    if out is None:
        out = torch.empty(*a.shape[:-1], Ca + Cb, device=a.device, dtype=BF16)
    rc = lib.imagd_concat_add_bf16(a.data_ptr(), lda, Ca, _ptr(res_a), _rows2d(res_a)[2] if res_a is not None else 0,
                                   _ptr(b), ldb, Cb, _ptr(res_b), _rows2d(res_b)[2] if res_b is not None else 0,
                                   out.data_ptr(), _rows2d(out)[2], rows, _stream())
    _lib.check(rc, "imagd_concat_add_bf16")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    NB, H, W, C = x.shape
    assert x.is_contiguous() and x.dtype == BF16
    out = torch.empty(NB, 2 * H, 2 * W, C, device=x.device, dtype=BF16)
    _lib.check(lib.imagd_upsample2x_bf16(x.data_ptr(), out.data_ptr(), NB, H, W, C, _stream()), "imagd_upsample2x_bf16")
    return out


def im2col3x3_s2(x: torch.Tensor, pad_lo: int = 1) -> torch.Tensor:
    """Stride-2 3x3 patches, tap-major. pad_lo = 1: symmetric padding 1 (UNet Downsample2D); pad_lo = 0: the VAE
    encoder's (0,1,0,1) right/bottom-only padding."""
    lib = _lib.load()
    NB, H, W, C = x.shape
    assert x.is_contiguous() and x.dtype == BF16
    out = torch.empty(NB, H // 2, W // 2, 9 * C, device=x.device, dtype=BF16)
    if pad_lo == 1:
        _lib.check(lib.imagd_im2col3x3_s2_bf16(x.data_ptr(), out.data_ptr(), NB, H, W, C, _stream()),
                   "imagd_im2col3x3_s2_bf16")
    else:
        _lib.check(lib.imagd_im2col3x3_s2_pad_bf16(x.data_ptr(), out.data_ptr(), NB, H, W, C, int(pad_lo), _stream()),
                   "imagd_im2col3x3_s2_pad_bf16")
    return out


def softmax_rows(s: torch.Tensor, scale: float = 1.0, *, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(scale * s) over the last dim: fp32 [rows, cols] -> bf16 [rows, cols]."""
    lib = _lib.load()
    assert s.dtype == torch.float32 and s.dim() == 2 and s.stride(1) == 1
    rows, cols = s.shape
    if out is None:
        out = torch.empty(rows, cols, device=s.device, dtype=BF16)
    _lib.check(lib.imagd_softmax_rows(s.data_ptr(), s.stride(0), out.data_ptr(), out.stride(0), rows, cols, float(scale),
                                      _stream()), "imagd_softmax_rows")
    return out


def conv3x3_direct(x: torch.Tensor, w: torch.Tensor, bias, *, stride: int = 1, act: int = ACT_NONE,
                   out_nchw_f32: bool = False, add: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    NB, H, W, Cin = x.shape
    assert x.is_contiguous() and x.dtype == BF16 and w.dtype == BF16
    Cout = w.shape[0]
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    if out is None:
        if out_nchw_f32:
            out = torch.empty(NB, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
        else:
            out = torch.empty(NB, Ho, Wo, Cout, device=x.device, dtype=BF16)
    rc = lib.imagd_conv3x3_direct_bf16(x.data_ptr(), NB, H, W, Cin, w.data_ptr(), _ptr(bias), out.data_ptr(), Cout,
                                       stride, act, 1 if out_nchw_f32 else 0, _ptr(add), _stream())
    _lib.check(rc, "imagd_conv3x3_direct_bf16")
    return out


def nchw_f32_to_nhwc_bf16(x: torch.Tensor, cpad: Optional[int] = None, *, repeat: int = 1, out=None) -> torch.Tensor:
    lib = _lib.load()
    NB, C, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    cpad = cpad or C
    if out is None:
        out = torch.empty(NB * repeat, H, W, cpad, device=x.device, dtype=BF16)
    _lib.check(lib.imagd_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), NB, C, H, W, cpad, repeat, _stream()),
               "imagd_nchw_f32_to_nhwc_bf16")
    return out


def timestep_embedding(timesteps: torch.Tensor, step_ptr: Optional[torch.Tensor], NB: int, dim: int, *,
                       out=None) -> torch.Tensor:
    lib = _lib.load()
    assert timesteps.dtype == torch.float32
    if out is None:
        out = torch.empty(NB, dim, device=timesteps.device, dtype=torch.float32)
    _lib.check(lib.imagd_timestep_embedding(timesteps.data_ptr(), _ptr(step_ptr), out.data_ptr(), NB, dim, _stream()),
               "imagd_timestep_embedding")
    return out


def linear_small_m(x: torch.Tensor, w: torch.Tensor, bias, *, act_in: int = ACT_NONE, act_out: int = ACT_NONE,
                   out=None) -> torch.Tensor:
    lib = _lib.load()
    M, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.float32 and w.dtype == BF16 and x.stride(1) == 1 and w.shape[1] == K
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    rc = lib.imagd_linear_small_m(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), out.data_ptr(),
                                  out.stride(0), M, N, K, act_in, act_out, _stream())
    _lib.check(rc, "imagd_linear_small_m")
    return out


def cfg_ddim_step(eps_cond: torch.Tensor, eps_uncond: Optional[torch.Tensor], guidance: float, latents: torch.Tensor,
                  coef: torch.Tensor, step_ptr: torch.Tensor, *, mask=None, image_latents=None, noise=None,
                  blend_coef=None) -> torch.Tensor:
    """In-place on `latents` (fp32 NCHW). step_ptr: int32[2] device tensor {step, scratch}."""
    lib = _lib.load()
    NB, C, H, W = latents.shape
    assert latents.dtype == torch.float32 and latents.is_contiguous() and eps_cond.dtype == torch.float32
    assert step_ptr.dtype == torch.int32 and step_ptr.numel() >= 2
    rc = lib.imagd_cfg_ddim_step(eps_cond.data_ptr(), _ptr(eps_uncond), float(guidance), latents.data_ptr(),
                                 coef.data_ptr(), step_ptr.data_ptr(), _ptr(mask), _ptr(image_latents), _ptr(noise),
                                 _ptr(blend_coef), NB, C, H * W, _stream())
    _lib.check(rc, "imagd_cfg_ddim_step")
    return latents


def embed_tokens(ids: torch.Tensor, tok: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """ids int64 [B, T]; tok bf16 [V, C]; pos bf16 [>= T, C] -> bf16 [B, T, C] = tok[ids] + pos[:T]."""
    lib = _lib.load()
    B, T = ids.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous() and tok.dtype == BF16 and pos.dtype == BF16
    assert tok.is_contiguous() and pos.is_contiguous() and pos.shape[0] >= T and pos.shape[1] == tok.shape[1]
    out = torch.empty(B, T, tok.shape[1], device=ids.device, dtype=BF16)
    _lib.check(lib.imagd_embed_tokens_bf16(ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), out.data_ptr(), B * T, T,
                                           tok.shape[1], tok.shape[0], _stream()), "imagd_embed_tokens_bf16")
    return out


def patchify(x: torch.Tensor, patch: int, kpad: int) -> torch.Tensor:
    """fp32 [B, 3, H, W] -> bf16 [B * (H/patch) * (W/patch), kpad] patch rows (column = (c*patch + iy)*patch + ix)."""
    lib = _lib.load()
    B, C, H, W = x.shape
    assert C == 3 and x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(B * (H // patch) * (W // patch), kpad, device=x.device, dtype=BF16)
    _lib.check(lib.imagd_patchify_bf16(x.data_ptr(), out.data_ptr(), B, H, W, patch, kpad, _stream()), "imagd_patchify_bf16")
    return out


def broadcast_row(vec: torch.Tensor, out: torch.Tensor, row: int) -> torch.Tensor:
    """out[b, row, :] = vec for every sample of out [B, rows, C] (bf16, contiguous)."""
    lib = _lib.load()
    B, R, C = out.shape
    assert out.is_contiguous() and out.dtype == BF16 and vec.dtype == BF16 and vec.numel() == C
    _lib.check(lib.imagd_broadcast_row_bf16(vec.data_ptr(), out.data_ptr(), B, R, row, C, _stream()),
               "imagd_broadcast_row_bf16")
    return out
