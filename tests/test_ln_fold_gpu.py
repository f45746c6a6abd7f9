"""GPU (r2-prep, not yet run on hardware): LayerNorm folded into the consuming GEMM — kernel level (producer row
statistics, LINEAR and GEGLU consumers through the C ABI vs the fp32 oracle ops) and model level (UNet eps with the fold
on vs off, and vs the oracle within the calibrated tolerance of test_unet_gpu)."""
import pytest
import torch

from conftest import rel_l2
from oracle import ops_ref

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (64, 1280)])
def test_producer_row_statistics(cuda_device, M, C):
    from imagdressing_b200 import ops

    a, w = _rand((M, C), cuda_device, 1).to(BF), _rand((C, C), cuda_device, 2, C ** -0.5).to(BF)
    b, res = _rand((C,), cuda_device, 3), (_rand((M, C), cuda_device, 4) + 2.0).to(BF)
    parts = ops.gemm_tile_count_n(M, C, C)
    stats = torch.full((M, parts, 2), float("nan"), device=cuda_device)
    y = ops.gemm(a, w, bias=b, residual=res, stats_out=stats)
    assert rel_l2(y, ops_ref.gemm_ref(a, w, b, residual=res)) < 1e-2
    tot = stats.sum(1)
    yf = y.float()
    assert torch.isfinite(stats).all()
    assert torch.allclose(tot[:, 0], yf.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[:, 1], (yf * yf).sum(1), rtol=1e-4, atol=1e-2)
    again = torch.empty_like(stats)
    ops.gemm(a, w, bias=b, residual=res, stats_out=again)
    assert torch.equal(stats, again)  # deterministic


@pytest.mark.parametrize("M,C,N", [(4096, 320, 960), (1000, 640, 640), (64, 1280, 3840)])
def test_consumer_linear_and_geglu(cuda_device, M, C, N):
    from imagdressing_b200 import modeling, ops

    dev = cuda_device
    a, w0 = _rand((M, C), dev, 5).to(BF), _rand((C, C), dev, 6, C ** -0.5).to(BF)
    res = (_rand((M, C), dev, 7) + 1.5).to(BF)
    parts = ops.gemm_tile_count_n(M, C, C)
    stats = torch.zeros(M, parts, 2, device=dev)
    x = ops.gemm(a, w0, residual=res, stats_out=stats)  # the raw residual stream + its row statistics
    gamma, beta = 1 + 0.2 * _rand((C,), dev, 8), 0.1 * _rand((C,), dev, 9)
    ln = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    # LINEAR consumer (fused QKV shape)
    w = _rand((N, C), dev, 10, C ** -0.5)
    wp, bp, cs = modeling.fold_layernorm(w, None, gamma, beta)
    out = ops.gemm(x, wp, bias=bp, ln=ops.LnFold(stats, parts, C, 1e-5, cs))
    assert rel_l2(out, ln @ w.T) < 1e-2
    # GEGLU consumer
    wg, bg = _rand((8 * C, C), dev, 11, C ** -0.5), _rand((8 * C,), dev, 12, 0.1)
    wf = wg * gamma[None, :]
    bf = bg + wg @ beta
    w1f, b1f = modeling.pack_geglu(wf, bf)
    out_g = ops.gemm(x, w1f, bias=b1f, act=ops.ACT_GEGLU, ln=ops.LnFold(stats, parts, C, 1e-5, w1f.float().sum(1).contiguous()))
    assert rel_l2(out_g, ops_ref.geglu_ref(ln, wg, bg)) < 1e-2


@torch.no_grad()
def test_unet_eps_with_fold_matches_plain_and_oracle(cuda_device):
    from imagdressing_b200 import modeling
    from test_unet_gpu import build_pair

    dev = cuda_device
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(2, 4, 32, 32, generator=g).to(dev)
    text = torch.randn(2, 77, 768, generator=g).to(dev)
    o, p = build_pair(dev, seed=0, with_ref=True)
    t = torch.tensor(981, device=dev)
    ref = torch.cat([o(lat[i:i + 1], t, text[i:i + 1])[0] for i in range(2)])
    plain = p(lat, t, text, return_dict=False)[0]
    old = modeling.FOLD_LN
    modeling.FOLD_LN = True
    try:
        p.invalidate_packed()
        folded = p(lat, t, text, return_dict=False)[0]
        again = p(lat, t, text, return_dict=False)[0]
    finally:
        modeling.FOLD_LN = old
        p.invalidate_packed()
    e_plain, e_fold = rel_l2(plain, ref), rel_l2(folded, ref)
    print(f"eps rel-L2 vs oracle: plain {e_plain:.4f}, LayerNorm folded {e_fold:.4f}; folded vs plain {rel_l2(folded, plain):.4f}")
    assert e_fold < max(1.5 * e_plain, 2e-2) and e_fold < 5e-2
    assert torch.equal(folded, again)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280),
                                             (3, 12, 9, 128, 64)])
def test_upconv_phase_kernel(cuda_device, NB, H, W, Cin, Cout):
    """imagd_upconv3x3_bf16 (r2-prep): nearest-2x upsample + conv3x3 as phase convs vs the oracle conv of the upsampled input."""
    from imagdressing_b200 import modeling, ops

    x = _rand((NB, H, W, Cin), cuda_device, 21).to(BF)
    w = _rand((Cout, Cin, 3, 3), cuda_device, 22, (9 * Cin) ** -0.5).to(BF)
    b = _rand((Cout,), cuda_device, 23)
    up = x.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()
    ref = ops_ref.conv3x3_ref(up, w, b)
    out = ops.upconv3x3(x, modeling.pack_upconv3x3(w), bias=b)
    assert out.shape == (NB, 2 * H, 2 * W, Cout)
    assert rel_l2(out, ref) < 1e-2
    assert rel_l2(out, ops.conv3x3(ops.upsample2x(x), ops_ref.conv3x3_pack(w), bias=b)) < 1e-2
