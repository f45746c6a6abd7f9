"""CPU: the LayerNorm-fold protocol (producer row statistics -> consumer epilogue) emulated in torch with the product's
own host packing (`modeling.fold_layernorm`, the block's folded GEGLU pack) against LN -> Linear in fp64. Checks the
algebra, the choice of statistics (bf16-rounded outputs, per-N-tile partials, E[x^2] - mean^2 in fp32) and the packing —
everything of the feature except the CUDA code itself (GPU: tests/test_ln_fold_gpu.py)."""
import pytest
import torch

from oracle import ops_ref

BF = torch.bfloat16


def producer(x_in, w, b, res, tile):
    """What the producing GEMM writes: bf16 outputs and, per row and N tile, {sum, sum of squares} of the rounded values."""
    y = (x_in.float() @ w.float().T + b + res.float()).to(BF)
    N = y.shape[1]
    parts = [y[:, c:c + tile].float() for c in range(0, N, tile)]
    stats = torch.stack([torch.stack([p.sum(1), (p * p).sum(1)], -1) for p in parts], 1)  # [M, parts, 2] fp32
    return y, stats


def consumer(y, stats, wp, bp, colsum, eps):
    """The consuming GEMM's epilogue: rstd * (acc - mean * colsum) + b' with acc = y W'^T in fp32."""
    C = y.shape[1]
    s = stats.sum(1)
    mean = s[:, 0] / C
    var = (s[:, 1] / C - mean * mean).clamp_min(0)
    rstd = torch.rsqrt(var + eps)
    acc = y.float() @ wp.float().T
    return rstd[:, None] * (acc - mean[:, None] * colsum[None, :]) + bp[None, :]


@pytest.mark.parametrize("C,N,tile,mean_shift", [(320, 960, 160, 0.0), (640, 640, 64, 3.0), (1280, 1280, 256, 10.0)])
def test_folded_linear_matches_layernorm_then_linear(C, N, tile, mean_shift):
    from imagdressing_b200.modeling import fold_layernorm

    g = torch.Generator().manual_seed(C)
    r = lambda *s: torch.randn(*s, generator=g)
    M = 96
    x_in, w_prod, b_prod = r(M, C).to(BF), (r(C, C) * C ** -0.5).to(BF), r(C) * 0.1
    res = (r(M, C) + mean_shift).to(BF)
    gamma, beta = 1 + 0.2 * r(C), 0.1 * r(C)
    w, b = r(N, C) * C ** -0.5, 0.1 * r(N)
    y, stats = producer(x_in, w_prod, b_prod, res, tile)
    wp, bp, colsum = fold_layernorm(w, b, gamma, beta)
    assert wp.dtype == BF and bp.dtype == torch.float32 and colsum.shape == (N,)
    out = consumer(y, stats, wp, bp, colsum, 1e-5)
    yd = y.double()
    ln = torch.nn.functional.layer_norm(yd, (C,), gamma.double(), beta.double(), 1e-5)
    ref = ln @ w.double().T + b.double()
    today = torch.nn.functional.layer_norm(y.float(), (C,), gamma, beta, 1e-5).to(BF).float() @ w.to(BF).float().T + b
    e_fold = float((out.double() - ref).norm() / ref.norm())
    e_today = float((today.double() - ref).norm() / ref.norm())
    assert e_fold < 4e-3 and e_fold < 1.5 * e_today, (e_fold, e_today)


def test_block_packs_folded_geglu_consistently():
    from imagdressing_b200 import modeling

    torch.manual_seed(0)
    old = modeling.FOLD_LN
    modeling.FOLD_LN = True
    try:
        blk = modeling.BasicTransformerBlock(64, 4, 32)
        torch.nn.init.normal_(blk.norm3.weight, 1.0, 0.2)
        torch.nn.init.normal_(blk.norm3.bias, 0.0, 0.2)
        pk = blk._packed()
    finally:
        modeling.FOLD_LN = old
    proj = blk.ff.net[0].proj
    C = 64
    x = torch.randn(40, C).to(BF)
    # consumer emulation on the packed tensors, GEGLU applied on the packed layout (64 value | 64 gate per 128 rows)
    stats = torch.stack([x.float().sum(1), (x.float() ** 2).sum(1)], -1)[:, None, :]
    with torch.no_grad():
        z = consumer(x, stats, pk["w1f"], pk["b1f"], pk["c1f"], blk.norm3.eps).view(40, -1, 2, 64)
    out = (z[:, :, 0] * torch.nn.functional.gelu(z[:, :, 1])).reshape(40, -1)
    ln = torch.nn.functional.layer_norm(x.float(), (C,), blk.norm3.weight, blk.norm3.bias, blk.norm3.eps)
    ref = ops_ref.geglu_ref(ln.detach(), proj.weight.detach(), proj.bias.detach())
    assert float((out - ref).norm() / ref.norm()) < 5e-3
    # the unfolded pack is untouched by the flag
    w1, b1 = modeling.pack_geglu(proj.weight, proj.bias)
    assert torch.equal(pk["w1"], w1) and torch.equal(pk["b1"], b1)


def test_cache_processor_declines_the_fold_and_foreign_processors_too():
    from adapter.attention_processor import CacheAttnProcessor2_0, CAttnProcessor2_0, RefSAttnProcessor2_0
    from imagdressing_b200 import modeling

    assert RefSAttnProcessor2_0("n", 64)._accepts_ln_fold and CAttnProcessor2_0("n", 64, 32)._accepts_ln_fold
    assert not CacheAttnProcessor2_0()._accepts_ln_fold  # the garment tap IS the LayerNorm output
    old = modeling.FOLD_LN
    modeling.FOLD_LN = True
    try:
        blk = modeling.BasicTransformerBlock(64, 4, 32)
        assert blk._can_fold()
        blk.attn1.set_processor(CacheAttnProcessor2_0())
        assert not blk._can_fold()
        blk.attn1.set_processor(lambda attn, hs, **kw: hs)  # a foreign callable: keeps the plain LayerNorm path
        assert not blk._can_fold()
    finally:
        modeling.FOLD_LN = old
