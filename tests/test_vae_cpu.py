"""CPU: host wiring of the kernel-backed AutoencoderKL (imagdressing_b200/vae.py) with the kernel wrappers emulated in torch
against the oracle restatement (oracle/vae.py): weight packing (quant_conv folded into conv_out, post_quant_conv as a centre
tap, padded thin ends), the (0,1,0,1) downsample, phase-conv upsample, the GEMM-softmax-GEMM mid attention with the value
bias folded into the out projection, state_dict compatibility incl. the deprecated attention key names."""
import pytest
import torch

import emulated_ops
from oracle import unet as ou
from oracle import vae as ov

CFG = dict(block_out_channels=(64, 128, 128, 128), norm_num_groups=32)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture
def emu(monkeypatch):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import vae

    return vae


@torch.no_grad()
def test_encode_decode_match_the_oracle(emu):
    o = ov.AutoencoderKL(**CFG).eval()
    ou.init_synthetic_(o, 3)
    p = emu.AutoencoderKL(**CFG).eval()
    assert set(p.state_dict()) == set(o.state_dict())
    p.load_state_dict(o.state_dict())
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    do, dp = o.encode(img).latent_dist, p.encode(img).latent_dist
    assert dp.mean.shape == (2, 4, 8, 6)
    assert rel(dp.mean, do.mean) < 2e-2 and rel(dp.logvar, do.logvar) < 2e-2
    z = torch.randn(2, 4, 8, 6, generator=g)
    xo, xp = o.decode(z)[0], p.decode(z, return_dict=False)[0]
    assert xp.shape == (2, 3, 64, 48) and rel(xp, xo) < 2e-2
    # sample(): mean + std * noise with the caller's generator (inpainting pipeline's _encode_vae_image)
    s1 = dp.sample(torch.Generator().manual_seed(5))
    s2 = dp.mean + dp.std * torch.randn(dp.mean.shape, generator=torch.Generator().manual_seed(5))
    assert torch.allclose(s1, s2) and torch.equal(dp.mode(), dp.mean)
    # deprecated attention key names (pre-0.18 VAE checkpoints: query / key / value / proj_attn, 1x1-conv shaped)
    old = {}
    for k, v in o.state_dict().items():
        for new, dep in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if f"attentions.0.{new}." in k:
                k = k.replace(f"attentions.0.{new}.", f"attentions.0.{dep}.")
                v = v[:, :, None, None] if v.dim() == 2 else v
        old[k] = v
    p2 = emu.AutoencoderKL(**CFG).eval()
    p2.load_state_dict(old)
    assert rel(p2.decode(z, return_dict=False)[0], xo) < 2e-2


@torch.no_grad()
def test_pipeline_edges_use_the_vae(emu, monkeypatch):
    """ref_image -> vae.encode(...).latent_dist.mean * 0.18215 and latents -> vae.decode(latents / scaling_factor) -> PIL,
    the two VAE call sites of the reference pipeline (IMAGDressing_v1_pipeline.py:454-458,544)."""
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline import IMAGDressing_v1
    from oracle.pipeline import sample_one
    from test_pipelines_cpu import STEPS, build, common, eager, inputs

    from imagdressing_b200 import modeling

    (o, ro, _), (p, rp, _), sched = build(modeling)
    vo = ov.AutoencoderKL(**CFG).eval()
    ou.init_synthetic_(vo, 3)
    vp = emu.AutoencoderKL(**CFG).eval()
    vp.load_state_dict(vo.state_dict())
    pipe = eager(IMAGDressing_v1(vae=vp, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, image_encoder=None,
                                 ImgProj=None, scheduler=sched, safety_checker=None, feature_extractor=None))
    assert pipe.vae_scale_factor == 8
    x = inputs(80)
    g = torch.Generator().manual_seed(1)
    ref_image = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    kw = common(x)
    kw.update(ref_image=ref_image, ref_image_latents=None, output_type="pil")
    images = pipe(guidance_scale=7.5, **kw).images
    assert len(images) == 1 and images[0].size == (128, 128)
    garment = vo.encode(ref_image).latent_dist.mean * 0.18215
    lat = sample_one(o, ro, x["latents"], x["prompt"], x["negative"], x["gtok"], garment, 7.5, STEPS)
    want = (vo.decode(lat / 0.18215)[0] / 2 + 0.5).clamp(0, 1)
    got = torch.from_numpy(__import__("numpy").array(images[0])).float().permute(2, 0, 1)[None] / 255.0
    assert rel(got, want) < 5e-2
