"""CPU: the training-step oracle (oracle/train_step.py) on a tiny SD1.5-shaped configuration —
which parameters receive gradients, known-answer values of the SNR weighting, and a finite-difference check that the
autograd graph through garment pass -> cached taps -> hybrid attention -> loss is complete."""
import math

import pytest
import torch

from oracle import processors as op
from oracle import train_step as ts
from oracle import unet as ou
from oracle.ddim import DDIMOracle

CFG = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)


def build(dtype=torch.float32):
    torch.manual_seed(0)
    unet, ref = ou.UNet2DConditionModel(**CFG), ou.UNet2DConditionModel(**CFG)
    ou.init_synthetic_(unet, 0)
    ou.init_synthetic_(ref, 1)
    proj = op.Resampler(dim=64, depth=1, dim_head=16, heads=4, num_queries=4, embedding_dim=48, output_dim=64, ff_mult=2)
    adapters = ts.install_training_processors(unet, ref)
    params = ts.set_trainable(unet, ref, proj, adapters)
    for m in (unet, ref, proj):
        m.to(dtype)
    return unet, ref, proj, adapters, params


def batch(B=2, dtype=torch.float32):
    g = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=g).to(dtype)
    return dict(latents=r(B, 4, 16, 16), ref_latents=r(B, 4, 16, 16), clip_image_embeddings=r(B, 9, 48),
                encoder_hidden_states=r(B, 7, 64), noise=r(B, 4, 16, 16), timesteps=torch.tensor([981, 40][:B]))


def test_snr_known_answers():
    s = DDIMOracle()
    t = torch.tensor([0, 500, 999])
    snr = ts.compute_snr(s.alphas_cumprod, t)
    a = s.alphas_cumprod[t]
    assert torch.allclose(snr, a / (1 - a), rtol=1e-5)
    assert snr[0] > 1000 and snr[2] < 0.01  # almost clean at t = 0, almost pure noise at t = 999
    pred, tgt = torch.zeros(3, 4, 2, 2), torch.ones(3, 4, 2, 2)
    assert float(ts.training_loss(pred, tgt)) == pytest.approx(1.0)
    # min-SNR-gamma: weight = min(snr, gamma) / snr -> 1 where snr <= gamma, gamma / snr above
    w = torch.minimum(snr, torch.tensor(5.0)) / snr
    assert float(ts.training_loss(pred, tgt, s.alphas_cumprod, t, snr_gamma=5.0)) == pytest.approx(float(w.mean()), rel=1e-6)


def test_processor_setup_mirrors_train_py():
    unet, ref, proj, adapters, params = build()
    names = list(unet.attn_processors)
    assert len(adapters) == 32 and len(names) == 32
    for n, p in unet.attn_processors.items():
        if n.endswith("attn1.processor"):
            layer = dict(unet.named_modules())[n.rsplit(".processor", 1)[0]]
            assert torch.equal(p.to_k_ref.weight, layer.to_k.weight) and torch.equal(p.to_v_ref.weight, layer.to_v.weight)
            assert p.to_k_ref.weight.requires_grad  # adapter modules trainable ...
            assert not layer.to_k.weight.requires_grad  # ... the denoising UNet itself frozen
        else:
            assert isinstance(p, op.CAttnProcessor)
    n_adapter = sum(p.numel() for p in adapters.parameters())
    n_hidden = [ts.hidden_size_of(n, CFG["block_out_channels"]) for n in names if n.endswith("attn1.processor")]
    assert n_adapter == sum(2 * h * h for h in n_hidden)
    assert len(params) == len(list(proj.parameters())) + len(list(ref.parameters())) + len(list(adapters.parameters()))


def test_train_step_gradient_routing():
    unet, ref, proj, adapters, params = build()
    loss = ts.train_step(unet, ref, proj, DDIMOracle(), snr_gamma=5.0, **batch())
    assert math.isfinite(float(loss)) and float(loss) > 0
    # frozen: nothing in the denoising UNet proper accumulates a gradient
    frozen = [n for n, p in unet.named_parameters() if ".processor." not in n]
    assert all(dict(unet.named_parameters())[n].grad is None for n in frozen)
    # trainable and reached: every to_k_ref / to_v_ref, the image projection, the garment UNet up to its last tap
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0
               for p in adapters.parameters())
    assert all(p.grad is not None for p in proj.parameters())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    assert rg["conv_in.weight"] is not None and rg["down_blocks.0.attentions.0.transformer_blocks.0.norm1.weight"] is not None
    # the garment UNet's own output is discarded (train.py:259): whatever follows the last self-attention tap gets no
    # gradient, and neither do the attn2 caches' consumers (only attn1 taps are read, attention_processor.py:598)
    assert rg["conv_out.weight"] is None and rg["conv_norm_out.weight"] is None
    last = "up_blocks.3.attentions.2.transformer_blocks.0"
    assert rg[last + ".norm1.weight"] is not None          # feeds the last tap
    assert rg[last + ".attn1.to_q.weight"] is None         # the garment UNet's own attention after the tap: unused
    assert rg[last + ".ff.net.2.weight"] is None


def test_train_step_directional_derivative_matches_autograd():
    """fp64: d loss / d eps along a random direction in (to_k_ref of one layer, garment conv_in) == <grad, direction>."""
    unet, ref, proj, adapters, params = build(torch.float64)
    b = batch(dtype=torch.float64)
    sched = DDIMOracle()
    sched.alphas_cumprod = sched.alphas_cumprod.double()

    def loss_only():
        noisy = sched.add_noise(b["latents"], b["noise"], b["timesteps"])
        pred = ts.sd_forward(unet, ref, proj, b["encoder_hidden_states"], noisy, b["ref_latents"],
                             b["clip_image_embeddings"], b["timesteps"])
        return ((pred - b["noise"]) ** 2).mean()  # training_loss casts to fp32 like train.py:576; keep fp64 here

    targets = [unet.attn_processors["down_blocks.1.attentions.0.transformer_blocks.0.attn1.processor"].to_k_ref.weight,
               ref.conv_in.weight, proj.proj_out.weight]
    loss_only().backward()
    g = torch.Generator().manual_seed(9)
    dirs = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in targets]
    analytic = sum(float((t.grad * d).sum()) for t, d in zip(targets, dirs))
    eps = 1e-6
    with torch.no_grad():
        for t, d in zip(targets, dirs):
            t.add_(eps * d)
        lp = float(loss_only())
        for t, d in zip(targets, dirs):
            t.sub_(2 * eps * d)
        lm = float(loss_only())
        for t, d in zip(targets, dirs):
            t.add_(eps * d)
    numeric = (lp - lm) / (2 * eps)
    assert numeric == pytest.approx(analytic, rel=1e-5, abs=1e-10)
