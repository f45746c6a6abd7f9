"""CPU: the host-side wiring of the LayerNorm fold — Transformer2DModel / BasicTransformerBlock / attention_forward run
with the kernel wrappers replaced by torch emulations of what the kernels compute (GEMM epilogues incl. row statistics
and the folded consumer formula, LayerNorm, two-stream attention). The folded run must match the plain run: catches
mis-wired statistics buffers, gamma / beta, residuals and the block <-> processor handshake without a GPU."""
import math

import pytest
import torch

BF = torch.bfloat16
TILE = 160  # emulated producer N tile


class FakeStream:
    def __init__(self, k, v, length, sample_rows=0, broadcast=False, n_query_samples=1 << 30, out_scale=1.0):
        self.k, self.v, self.length, self.sample_rows = k, v, length, sample_rows or length
        self.broadcast, self.n_query_samples, self.out_scale = broadcast, n_query_samples, out_scale


def fake_gemm(a, w, *, out=None, bias=None, rowvec=None, rows_per_group=0, residual=None, act=0, alpha=1.0,
              out_fp32=False, stats_out=None, ln=None):
    from imagdressing_b200.ops import ACT_GEGLU

    lead = a.shape[:-1]
    a2 = a.reshape(-1, a.shape[-1]).float()
    acc = alpha * (a2 @ w.float().T)
    if ln is not None:
        s = ln.stats[:, :ln.parts].sum(1)
        mean = s[:, 0] / ln.dim
        rstd = torch.rsqrt((s[:, 1] / ln.dim - mean * mean).clamp_min(0) + ln.eps)
        y = rstd[:, None] * (acc - mean[:, None] * ln.colsum[None, :]) + bias[None, :]
    else:
        y = acc + (bias[None, :] if bias is not None else 0)
    if act == ACT_GEGLU:
        z = y.view(y.shape[0], -1, 2, 64)
        y = (z[:, :, 0] * torch.nn.functional.gelu(z[:, :, 1])).reshape(y.shape[0], -1)
    if residual is not None:
        y = y + residual.reshape(-1, residual.shape[-1]).float()
    y = y.to(BF)
    if stats_out is not None:
        stats_out.zero_()
        for i, c in enumerate(range(0, y.shape[1], TILE)):
            part = y[:, c:c + TILE].float()
            stats_out[:, i, 0] = part.sum(1)
            stats_out[:, i, 1] = (part * part).sum(1)
    y = y.reshape(*lead, y.shape[-1])
    if out is not None:
        out.copy_(y)
        return out
    return y


def fake_layernorm(x, gamma, beta, eps=1e-5, *, out=None):
    return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).to(BF)


def fake_attention(q, B, Lq, heads, hd, s0, s1=None, *, sm_scale=None, out=None):
    C = heads * hd
    qf = q.float().reshape(B, Lq, heads, hd).transpose(1, 2)

    def stream(s, b):
        src = 0 if s.broadcast else b
        rows = slice(src * s.sample_rows, src * s.sample_rows + s.length)
        k = s.k[rows].float().reshape(s.length, heads, hd).transpose(0, 1)
        v = s.v[rows].float().reshape(s.length, heads, hd).transpose(0, 1)
        return torch.softmax(qf[b] @ k.transpose(1, 2) / math.sqrt(hd), -1) @ v  # [heads, Lq, hd]

    res = []
    for b in range(B):
        o = stream(s0, b)
        if s1 is not None and b < s1.n_query_samples:
            o = o + s1.out_scale * stream(s1, b)
        res.append(o.transpose(0, 1).reshape(Lq, C))
    return torch.stack(res).reshape(B * Lq, C).to(BF)


@pytest.fixture
def emulated(monkeypatch):
    from imagdressing_b200 import modeling, ops

    monkeypatch.setattr(ops, "gemm", fake_gemm)
    monkeypatch.setattr(ops, "layernorm", fake_layernorm)
    monkeypatch.setattr(ops, "attention", fake_attention)
    monkeypatch.setattr(ops, "kv_stream", lambda k, v, length, **kw: FakeStream(k, v, length, **kw))
    monkeypatch.setattr(ops, "gemm_tile_count_n", lambda M, N, K: (N + TILE - 1) // TILE)
    monkeypatch.setattr(ops, "groupnorm", lambda x, g, b, groups, eps, silu, out=None, ws=None: torch.nn.functional.group_norm(
        x.float().permute(0, 3, 1, 2), groups, g, b, eps).permute(0, 2, 3, 1).to(BF).contiguous())
    return modeling


@pytest.mark.parametrize("hybrid", [False, True])
def test_transformer_block_folded_equals_plain(emulated, hybrid):
    modeling = emulated
    from adapter.attention_processor import CAttnProcessor2_0, RefSAttnProcessor2_0

    torch.manual_seed(0)
    C, heads, NB, H, W = 320, 8, 2, 4, 4
    tr = modeling.Transformer2DModel(C, heads, 64, 32)
    for n, p in tr.named_parameters():
        torch.nn.init.normal_(p, 1.0 if "norm" in n and n.endswith("weight") else 0.0, 0.2 if "norm" in n else C ** -0.5)
    name = "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor"
    blk = tr.transformer_blocks[0]
    blk.attn1.set_processor(RefSAttnProcessor2_0(name, C, scale=0.8))
    blk.attn2.set_processor(CAttnProcessor2_0(name.replace("attn1", "attn2"), C, 64))
    x = torch.randn(NB, H, W, C).to(BF)
    ctx = torch.randn(NB, 7, 64).to(BF)
    kw = {"sa_hidden_states": {name: torch.randn(1, 16, C).to(BF)}, "ref_samples": 1} if hybrid else {}

    def run(fold):
        modeling.FOLD_LN = fold
        tr._invalidate()
        blk._invalidate()
        for a in (blk.attn1, blk.attn2):
            a.invalidate_packed()
            a.processor.invalidate_packed()
        return tr.run(x, ctx, dict(kw)).float()

    try:
        plain, folded = run(False), run(True)
        assert blk._can_fold() is True
    finally:
        modeling.FOLD_LN = False
    err = float((folded - plain).norm() / plain.norm())
    assert err < 1e-2, err  # both are bf16-chained; they differ only by where the roundings fall
    # and the fold really was taken: the statistics buffers exist and hold the sums of the block's streams
    assert hasattr(tr, "_row_stats") and hasattr(blk, "_row_stats")
    (s0, p0), = [v for k, v in tr._row_stats._bufs.items() if k[0] == "x0"]
    assert p0 == 2 and torch.isfinite(s0).all() and float(s0.abs().sum()) > 0
