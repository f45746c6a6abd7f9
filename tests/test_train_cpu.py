"""CPU: the training step's host side with the kernel wrappers replaced by torch emulations that implement
the backward kernels' EXPLICIT formulas (tests/emulated_ops.py): every autograd operator against torch autograd of the
reference operator, then the whole SDModel step (Resampler -> garment UNet taps -> hybrid denoising UNet -> MSE -> backward)
against the oracle (oracle/train_step.py, which restates /root/reference/train.py:255-281,338-379,573-605) on a tiny
SD1.5-shaped configuration: loss, which parameters receive gradients, and every gradient. The kernels' arithmetic is checked
on the GPU (tests/test_train_ops_gpu.py, tests/test_train_step_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

import emulated_ops
from oracle import processors as op
from oracle import train_step as ts
from oracle import unet as ou
from oracle.ddim import DDIMOracle

BF = torch.bfloat16
CFG = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


@pytest.fixture
def emu(monkeypatch):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import autograd as ag

    ag.clear_cache()
    return ag


def rb(*shape, seed=0, scale=1.0):
    """bf16-representable fp32 values."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF).float()


def leaves(*ts_):
    return [t.clone().requires_grad_(True) for t in ts_]


def check(got, want, tol=2e-2):
    for g, w in zip(got, want):
        assert g is not None and torch.isfinite(g.float()).all()
        assert rel(g, w) < tol, rel(g, w)


def test_linear_and_conv_operators(emu):
    ag = emu
    x, w, b, r = rb(2, 12, 32, seed=1), rb(48, 32, seed=2, scale=0.2), rb(48, seed=3), rb(2, 12, 48, seed=4)
    xl, wl, bl, rl = leaves(x, w, b, r)
    (F.linear(xl, wl, bl) + rl).square().sum().backward()
    xp, wp, bp, rp = leaves(x.to(BF), w.to(BF), b, r.to(BF))
    y = ag.linear(xp, wp, bp, rp)
    y.float().square().sum().backward()
    check([xp.grad, wp.grad, bp.grad, rp.grad], [xl.grad, wl.grad, bl.grad, rl.grad])

    # conv3x3 with bias + per-sample row vector + residual
    x, w, b = rb(2, 6, 5, 16, seed=5), rb(24, 16, 3, 3, seed=6, scale=0.1), rb(24, seed=7)
    tv, r = rb(2, 24, seed=8), rb(2, 6, 5, 24, seed=9)
    xl, wl, bl, tl, rl = leaves(x, w, b, tv, r)
    yr = F.conv2d(xl.permute(0, 3, 1, 2), wl, bl, padding=1) + tl[:, :, None, None]
    (yr.permute(0, 2, 3, 1) + rl).square().sum().backward()
    xp, wq, bp, tp, rp = leaves(x.to(BF), w, b, tv, r.to(BF))
    y = ag.conv3x3(xp, ag.pack_conv3x3(wq), bp, tp, rp)
    y.float().square().sum().backward()
    check([xp.grad, wq.grad, bp.grad, tp.grad, rp.grad], [xl.grad.to(BF), wl.grad, bl.grad, tl.grad, rl.grad])

    # conv_in (no input gradient) and conv_out (input gradient only, fp32 NCHW output)
    x, w, b = rb(2, 6, 5, 4, seed=10), rb(32, 4, 3, 3, seed=11, scale=0.2), rb(32, seed=12)
    wl, bl = leaves(w, b)
    F.conv2d(x.permute(0, 3, 1, 2), wl, bl, padding=1).square().sum().backward()
    wq, bp = leaves(w, b)
    ag.ConvIn.apply(x.to(BF), ag.pack_conv3x3(wq), bp).float().square().sum().backward()
    check([wq.grad, bp.grad], [wl.grad, bl.grad])
    x, w, b = rb(2, 6, 5, 32, seed=13), rb(4, 32, 3, 3, seed=14, scale=0.1), rb(4, seed=15)
    (xl,) = leaves(x)
    F.conv2d(xl.permute(0, 3, 1, 2), w, b, padding=1).square().sum().backward()
    (xp,) = leaves(x.to(BF))
    out = ag.ConvOut.apply(xp, ag.pack_conv3x3(w), b)
    assert out.shape == (2, 4, 6, 5) and out.dtype == torch.float32
    out.square().sum().backward()
    check([xp.grad], [xl.grad])

    # Downsample2D = stride-2 im2col + GEMM; Upsample2D = nearest 2x + conv
    x, w, b = rb(2, 8, 6, 16, seed=16), rb(16, 16, 3, 3, seed=17, scale=0.1), rb(16, seed=18)
    xl, wl, bl = leaves(x, w, b)
    F.conv2d(xl.permute(0, 3, 1, 2), wl, bl, stride=2, padding=1).square().sum().backward()
    xp, wq, bp = leaves(x.to(BF), w, b)
    ag.linear(ag.Im2colS2.apply(xp), ag.pack_conv3x3(wq), bp).float().square().sum().backward()
    check([xp.grad, wq.grad, bp.grad], [xl.grad, wl.grad, bl.grad])
    xl, wl = leaves(x, w)
    F.conv2d(F.interpolate(xl.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), wl, padding=1).square().sum().backward()
    xp, wq = leaves(x.to(BF), w)
    ag.conv3x3(ag.Upsample2x.apply(xp), ag.pack_conv3x3(wq)).float().square().sum().backward()
    check([xp.grad, wq.grad], [xl.grad, wl.grad])


def test_norm_activation_concat_operators(emu):
    ag = emu
    x = rb(2, 6, 32, seed=1) * 1.5 + 0.3
    for silu in (True, False):
        gn_o, gn_p = torch.nn.GroupNorm(8, 32), torch.nn.GroupNorm(8, 32)
        with torch.no_grad():
            gn_o.weight.copy_(1 + 0.1 * rb(32, seed=2)); gn_o.bias.copy_(0.1 * rb(32, seed=3))
        gn_p.load_state_dict(gn_o.state_dict())
        (xl,) = leaves(x)
        y = gn_o(xl.movedim(-1, 1))
        (F.silu(y) if silu else y).square().sum().backward()
        (xp,) = leaves(x.to(BF))
        ag.groupnorm(xp, gn_p, silu).float().square().sum().backward()
        check([xp.grad, gn_p.weight.grad, gn_p.bias.grad], [xl.grad, gn_o.weight.grad, gn_o.bias.grad])
    ln_o, ln_p = torch.nn.LayerNorm(32), torch.nn.LayerNorm(32)
    with torch.no_grad():
        ln_o.weight.copy_(1 + 0.1 * rb(32, seed=4)); ln_o.bias.copy_(0.1 * rb(32, seed=5))
    ln_p.load_state_dict(ln_o.state_dict())
    (xl,) = leaves(x)
    ln_o(xl).square().sum().backward()
    (xp,) = leaves(x.to(BF))
    ag.layernorm(xp, ln_p).float().square().sum().backward()
    check([xp.grad, ln_p.weight.grad, ln_p.bias.grad], [xl.grad, ln_o.weight.grad, ln_o.bias.grad])
    for fn_p, fn_o in ((ag.silu, F.silu), (ag.gelu, F.gelu)):
        (xl,) = leaves(x)
        fn_o(xl).square().sum().backward()
        (xp,) = leaves(x.to(BF))
        fn_p(xp).float().square().sum().backward()
        check([xp.grad], [xl.grad])
    h = rb(2, 6, 64, seed=6)
    (hl,) = leaves(h)
    v, g = hl.chunk(2, -1)
    (v * F.gelu(g)).square().sum().backward()
    (hp,) = leaves(h.to(BF))
    ag.Geglu.apply(hp).float().square().sum().backward()
    check([hp.grad], [hl.grad])
    a, b = rb(2, 3, 4, 16, seed=7), rb(2, 3, 4, 24, seed=8)
    al, bl = leaves(a, b)
    (torch.cat([al, bl], -1) * rb(2, 3, 4, 40, seed=9)).sum().backward()
    ap, bp = leaves(a.to(BF), b.to(BF))
    (ag.Concat.apply(ap, bp).float() * rb(2, 3, 4, 40, seed=9)).sum().backward()
    check([ap.grad, bp.grad], [al.grad, bl.grad], 1e-2)
    p, t = rb(2, 4, 5, 5, seed=10), rb(2, 4, 5, 5, seed=11)
    (pl,) = leaves(p)
    lo = F.mse_loss(pl, t)
    lo.backward()
    (pp,) = leaves(p)
    lp = ag.mse_loss(pp, t)
    lp.backward()
    assert abs(float(lp) - float(lo)) < 1e-6 and rel(pp.grad, pl.grad) < 1e-6


def _ref_attn(q, k0, v0, k1, v1, heads, w1):
    B, L, C = q.shape
    sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, C // heads).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k0), sp(v0))
    if k1 is not None:
        o = o + w1 * F.scaled_dot_product_attention(sp(q), sp(k1), sp(v1))
    return o.transpose(1, 2).reshape(B, L, C)


def test_attention_operator(emu):
    ag = emu
    B, L, C, heads = 2, 10, 32, 4
    qkv, kv1, kvc = rb(B, L, 3 * C, seed=1), rb(B, 7, 2 * C, seed=2), rb(B, 9, 2 * C, seed=3)
    d = rb(B, L, C, seed=4)
    # fused self-attention + garment stream
    ql, k1l = leaves(qkv, kv1)
    _ref_attn(ql[..., :C], ql[..., C:2 * C], ql[..., 2 * C:], k1l[..., :C], k1l[..., C:], heads, 0.8).mul(d).sum().backward()
    qp, k1p = leaves(qkv.to(BF), kv1.to(BF))
    ag.attention(qp, None, k1p, heads, 0.8).float().mul(d).sum().backward()
    check([qp.grad, k1p.grad], [ql.grad, k1l.grad])
    # cross-attention on a 7-token window of a 9-token context: tokens outside the window get zero gradient
    q = rb(B, L, C, seed=5)
    ql, kl = leaves(q, kvc)
    _ref_attn(ql, kl[:, :7, :C], kl[:, :7, C:], None, None, heads, 0.0).mul(d).sum().backward()
    qp, kp = leaves(q.to(BF), kvc.to(BF))
    ag.attention(qp, kp, None, heads, 1.0, 7).float().mul(d).sum().backward()
    check([qp.grad, kp.grad], [ql.grad, kl.grad])
    assert (kp.grad[:, 7:] == 0).all()


# ------------------------------------------------------------------------------------------------ whole step vs the oracle
def build_pair():
    from adapter.resampler import Resampler
    from imagdressing_b200 import modeling, train

    torch.manual_seed(0)
    o_unet, o_ref = ou.UNet2DConditionModel(**CFG), ou.UNet2DConditionModel(**CFG)
    ou.init_synthetic_(o_unet, 0)
    ou.init_synthetic_(o_ref, 1)
    rs_kw = dict(dim=64, depth=2, dim_head=16, heads=4, num_queries=4, embedding_dim=48, output_dim=64, ff_mult=2)
    o_proj = op.Resampler(**rs_kw)
    for m in (o_unet, o_ref, o_proj):  # bf16-representable weights on both sides
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(p.to(BF).float())
    o_ad = ts.install_training_processors(o_unet, o_ref)
    o_params = ts.set_trainable(o_unet, o_ref, o_proj, o_ad)

    p_unet, p_ref = modeling.UNet2DConditionModel(**CFG), modeling.UNet2DConditionModel(**CFG)
    p_proj = Resampler(**rs_kw)
    p_unet.load_state_dict({k: v for k, v in o_unet.state_dict().items() if "processor" not in k})
    p_ref.load_state_dict(o_ref.state_dict())
    p_proj.load_state_dict(o_proj.state_dict())
    p_ad = train.install_training_processors(p_unet, p_ref)
    p_ad.load_state_dict(o_ad.state_dict())
    train.set_trainable(p_unet, p_ref, p_proj, p_ad)
    return (o_unet, o_ref, o_proj, o_ad), (p_unet, p_ref, p_proj, p_ad)


def batch(B=2):
    r = lambda *s, seed: rb(*s, seed=seed)
    return dict(latents=r(B, 4, 16, 16, seed=1), ref_latents=r(B, 4, 16, 16, seed=2), clip_image_embeddings=r(B, 9, 48, seed=3),
                encoder_hidden_states=r(B, 7, 64, seed=4), noise=r(B, 4, 16, 16, seed=5), timesteps=torch.tensor([981, 40][:B]))


def test_sdmodel_step_matches_oracle(emu):
    from imagdressing_b200 import train
    from imagdressing_b200.scheduler import DDIMScheduler

    (o_unet, o_ref, o_proj, o_ad), (p_unet, p_ref, p_proj, p_ad) = build_pair()
    b = batch()
    loss_o = ts.train_step(o_unet, o_ref, o_proj, DDIMOracle(), **b)
    sd = train.SDModel(p_unet, p_ref, p_proj, p_ad)
    loss_p = train.train_step(sd, DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False), **b)
    assert abs(float(loss_p) - float(loss_o)) < 2e-2 * float(loss_o), (float(loss_p), float(loss_o))

    # the frozen denoising UNet has no gradients; adapters, garment UNet and projection do (train.py:368-379)
    for n, p in p_unet.named_parameters():
        assert (p.grad is not None) == ("processor" in n), n
    go = {**{"ref." + n: p.grad for n, p in o_ref.named_parameters()}, **{"proj." + n: p.grad for n, p in o_proj.named_parameters()},
          **{"ad." + n: p.grad for n, p in o_ad.named_parameters()}}
    gp = {**{"ref." + n: p.grad for n, p in p_ref.named_parameters()}, **{"proj." + n: p.grad for n, p in p_proj.named_parameters()},
          **{"ad." + n: p.grad for n, p in p_ad.named_parameters()}}
    assert set(go) == set(gp)
    num = den = 0.0
    worst = (0.0, "")
    for n, g in go.items():
        if g is None or float(g.norm()) == 0.0:  # parameters behind the garment UNet's last tap: no gradient in either
            assert gp[n] is None or float(gp[n].float().norm()) < 1e-6 * (1 + float(g.norm()) if g is not None else 1), n
            continue
        assert gp[n] is not None, n
        e = rel(gp[n], g)
        num += float((gp[n].float() - g).norm()) ** 2
        den += float(g.norm()) ** 2
        if g.numel() >= 64 and e > worst[0]:
            worst = (e, n)
        assert e < 0.15, (n, e)  # single small tensors (bf16 chain noise); the aggregate bound below is the parity statement
    assert (num / den) ** 0.5 < 3e-2, ((num / den) ** 0.5, worst)


def test_min_snr_gamma_loss_matches_oracle(emu):
    """train.py:579-596: loss = mean_b(mse_b * min(snr_b, gamma) / snr_b), forward value and gradient."""
    from imagdressing_b200 import train
    from imagdressing_b200.scheduler import DDIMScheduler

    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False)
    t = torch.tensor([5, 400, 990])
    pred, tgt = rb(3, 4, 6, 6, seed=1), rb(3, 4, 6, 6, seed=2)
    (pl,) = leaves(pred)
    lo = ts.training_loss(pl, tgt, DDIMOracle().alphas_cumprod, t, snr_gamma=5.0)
    lo.backward()
    (pp,) = leaves(pred)
    lp = train.training_loss(pp, tgt, sched, t, snr_gamma=5.0)
    lp.backward()
    assert abs(float(lp) - float(lo)) < 1e-6 * abs(float(lo)) + 1e-9 and rel(pp.grad, pl.grad) < 1e-5
    assert torch.allclose(train.compute_snr(sched, t), ts.compute_snr(DDIMOracle().alphas_cumprod, t), rtol=1e-5)


def test_flat_adamw_updates_views(emu):
    from imagdressing_b200 import train

    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)).to(BF)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    ref.load_state_dict({k: v.float() for k, v in lin.state_dict().items()})
    opt = train.FlatAdamW(lin.parameters(), lr=1e-2, weight_decay=0.1, bucket_bytes=16, step_fn=emulated_ops.adamw_step)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    assert len(opt._buckets) > 1
    x = rb(4, 5, seed=1)
    for _ in range(3):
        opt.zero_grad()
        ropt.zero_grad()
        lin(x.to(BF)).float().square().sum().backward()
        ref(x).square().sum().backward()
        opt.step()
        ropt.step()
    for p, q in zip(lin.parameters(), ref.parameters()):
        assert p.dtype == BF and rel(p, q) < 2e-2
        assert p.data_ptr() >= opt.param.data_ptr() and p.data_ptr() < opt.param.data_ptr() + opt.param.numel() * 2


def test_accumulation_window_and_schedule_through_the_whole_step(emu):
    """CPU twin of tests/test_train_step_gpu.py::test_gradient_accumulation_and_lr_schedule_on_the_kernels (train.py:606-608):
    two one-sample micro-steps with accumulation_steps=2 against one step on the batch of two; LRScheduler's rate reaches the
    update through the device-side scalar."""
    from imagdressing_b200 import train
    from imagdressing_b200.scheduler import DDIMScheduler

    _, (p_unet, p_ref, p_proj, p_ad) = build_pair()
    sd = train.SDModel(p_unet, p_ref, p_proj, p_ad)
    for m in (p_unet, p_ref, p_proj):
        m.to(BF)
    params = train.set_trainable(p_unet, p_ref, p_proj, p_ad)
    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False)
    b = batch()
    opt = train.FlatAdamW(params, lr=1e-3, weight_decay=1e-2, accumulation_steps=2)
    start = (opt.param.clone(), opt.master.clone())
    lrs = train.LRScheduler("constant_with_warmup", opt, num_warmup_steps=1)
    assert lrs.get_lr()[0] == 0.0 and float(opt.hyper[0]) == 0.0 and float(opt.hyper[3]) == 0.5

    def window():
        out = []
        for k in range(2):
            out.append(float(train.train_step(sd, sched, **{key: v[k:k + 1] for key, v in b.items()}, optimizer=opt)))
            assert opt._micro == (1 if k == 0 else 0)
        return out

    window()
    assert opt.t == 1 and torch.equal(opt.param, start[0])  # rate 0: nothing moves
    lrs.step()
    opt.reset_state()
    l_acc = window()
    g_acc, p_acc = opt.grad.float().clone() * 0.5, opt.param.float().clone()
    assert not torch.equal(opt.param, start[0])
    with torch.no_grad():
        opt.param.copy_(start[0])
        opt.master.copy_(start[1])
    opt.reset_state()
    opt.accum = 1
    opt.hyper[3:4].fill_(1.0)
    torch.autograd.graph.increment_version([opt.param, *opt.params])
    l_whole = float(train.train_step(sd, sched, **b, optimizer=opt))
    assert abs(0.5 * (l_acc[0] + l_acc[1]) - l_whole) < 2e-3 * abs(l_whole)
    assert rel(g_acc, opt.grad.float()) < 2e-2
    moved = (opt.param.float() - start[0].float()).abs().mean()
    assert float((p_acc - opt.param.float()).abs().mean()) < 0.1 * float(moved)
