"""CPU: the oracle's restatement of the reference denoising loop (oracle/pipeline.py) on a tiny SD1.5-shaped
configuration — BASELINE.json configs[0] plumbing (20 DDIM steps, CFG, garment pass at step 0) and the loop-level
properties the product's parity tests rely on: CFG algebra, the unconditional branch never sees the garment stream,
the ControlNet residual path, the inpaint blend."""
import pytest
import torch

from oracle import processors as op
from oracle import unet as ou
from oracle.pipeline import sample_one
from oracle.train_step import hidden_size_of

CFG = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
STEPS = 20


@pytest.fixture(scope="module")
def models():
    torch.manual_seed(0)
    unet, ref, cn = ou.UNet2DConditionModel(**CFG), ou.UNet2DConditionModel(**CFG), ou.ControlNetModel(**CFG)
    unet.set_attn_processor({n: (op.RefSAttnProcessor(n, hidden_size_of(n, CFG["block_out_channels"]), scale=1.0)
                                 if "attn1" in n else op.CAttnProcessor(n)) for n in unet.attn_processors})
    ref.set_attn_processor({n: op.CacheAttnProcessor() for n in ref.attn_processors})
    for m, s in ((unet, 0), (ref, 1), (cn, 2)):
        ou.init_synthetic_(m, s)
    return unet.eval(), ref.eval(), cn.eval()


def inputs(seed=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(latents=r(1, 4, 16, 16), prompt=r(1, 7, 64), negative=r(1, 7, 64), gtok=r(1, 4, 64),
                garment=r(1, 4, 16, 16), pose=torch.rand(1, 3, 128, 128, generator=g))


def run(models, x, guidance=7.5, **kw):
    unet, ref, _ = models
    return sample_one(unet, ref, x["latents"], x["prompt"], x["negative"], x["gtok"], x["garment"], guidance, STEPS, **kw)


def test_config0_plumbing_is_deterministic_and_finite(models):
    x = inputs()
    a, b = run(models, x), run(models, x)
    assert a.shape == (1, 4, 16, 16) and torch.isfinite(a).all()
    assert torch.equal(a, b)
    assert not torch.allclose(a, run(models, inputs(4)))


def test_cfg_algebra_and_garment_stream_only_on_the_conditional_branch(models):
    unet, ref, _ = models
    x = inputs()
    # guidance 1 -> eps = eps_cond: the negative prompt must not matter
    y = dict(x, negative=torch.randn(1, 7, 64, generator=torch.Generator().manual_seed(99)))
    assert torch.allclose(run(models, x, guidance=1.0), run(models, y, guidance=1.0), atol=1e-5)
    # guidance 0 -> eps = eps_uncond: neither the prompt nor the GARMENT may matter (reference :511-518 omits the kwarg)
    z = dict(x, prompt=y["negative"], garment=x["garment"] * -2.0, gtok=x["gtok"] + 1.0)
    assert torch.allclose(run(models, x, guidance=0.0), run(models, z, guidance=0.0), atol=1e-5)
    # with guidance the garment does matter, through the RefS scale
    base = run(models, x)
    assert not torch.allclose(base, run(models, dict(x, garment=x["garment"] * -2.0)), atol=1e-3)
    for p in unet.attn_processors.values():
        if hasattr(p, "to_k_ref"):
            p.scale = 0.0
    try:
        off = run(models, x)
        assert torch.allclose(off, run(models, dict(x, garment=x["garment"] * -2.0)), atol=1e-5)  # scale 0 == no garment
    finally:
        for p in unet.attn_processors.values():
            if hasattr(p, "to_k_ref"):
                p.scale = 1.0


def test_controlnet_residuals_and_inpaint_blend(models):
    unet, ref, cn = models
    x = inputs()
    base = run(models, x)
    with_cn = run(models, x, controlnet=cn, control_cond=x["pose"], control_scale=1.0)
    assert not torch.allclose(base, with_cn, atol=1e-3)
    # conditioning_scale 0 zeroes every residual: identical to the plain loop
    assert torch.allclose(base, run(models, x, controlnet=cn, control_cond=x["pose"], control_scale=0.0), atol=1e-5)
    # inpainting: where the mask is 0 the result is exactly the original image latents; where it is 1, the sampled ones
    g = torch.Generator().manual_seed(11)
    img = torch.randn(1, 4, 16, 16, generator=g)
    mask = torch.zeros(1, 1, 16, 16)
    mask[..., 4:12, 4:12] = 1.0
    out = run(models, x, controlnet=cn, control_cond=x["pose"], mask=mask, image_latents=img, noise=x["latents"])
    keep = (mask == 0).expand_as(out)
    assert torch.equal(out[keep], img[keep])
    assert not torch.allclose(out[~keep], img[~keep], atol=1e-2)
