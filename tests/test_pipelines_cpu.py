"""CPU: the four public pipeline classes (dressing_sd.pipelines...IMAGDressing_v1) and the DenoiseEngine's eager path
executed with the kernel wrappers emulated in torch (tests/emulated_ops.py) against the oracle loop (oracle/pipeline.py)
on a tiny SD1.5-shaped configuration: argument routing, CFG batching, garment features, ControlNet / IP-Adapter / inpaint
plumbing, per-image state refresh. The CUDA-graph replay path and the kernels are covered by tests/test_pipeline_gpu.py."""
import pytest
import torch

import emulated_ops
from oracle import processors as op
from oracle import unet as ou
from oracle.pipeline import sample_one
from oracle.train_step import hidden_size_of

CFG = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
BOC = CFG["block_out_channels"]
H = W = 16
STEPS = 4


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture(params=[False, True], ids=["plain", "ln_fold"])
def emu(monkeypatch, request):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import modeling

    monkeypatch.setattr(modeling, "FOLD_LN", request.param)
    return modeling


def build(modeling, kind="base"):
    from adapter.attention_processor import (CacheAttnProcessor2_0, CAttnProcessor2_0, LoraRefSAttnProcessor2_0,
                                             LoRAIPAttnProcessor2_0, RefSAttnProcessor2_0)
    from imagdressing_b200.scheduler import DDIMScheduler

    o, p = ou.UNet2DConditionModel(**CFG), modeling.UNet2DConditionModel(**CFG)
    hs = lambda n: hidden_size_of(n, BOC)
    if kind == "ipa":
        o.set_attn_processor({n: (op.LoraRefSAttnProcessor(n, hs(n), rank=4, lora_scale=0.2, scale=0.9) if "attn1" in n
                                  else op.LoRAIPAttnProcessor(hs(n), 64, rank=4, lora_scale=0.3, scale=0.8, num_tokens=4))
                              for n in o.attn_processors})
        p.set_attn_processor({n: (LoraRefSAttnProcessor2_0(n, hs(n), rank=4) if "attn1" in n
                                  else LoRAIPAttnProcessor2_0(hs(n), 64, rank=4, num_tokens=4)) for n in p.attn_processors})
    else:
        o.set_attn_processor({n: (op.RefSAttnProcessor(n, hs(n), scale=1.0) if "attn1" in n else op.CAttnProcessor(n))
                              for n in o.attn_processors})
        p.set_attn_processor({n: (RefSAttnProcessor2_0(n, hs(n)) if "attn1" in n else CAttnProcessor2_0(n, hs(n), 64))
                              for n in p.attn_processors})
    ro, rp = ou.UNet2DConditionModel(**CFG), modeling.UNet2DConditionModel(**CFG)
    ro.set_attn_processor({n: op.CacheAttnProcessor() for n in ro.attn_processors})
    rp.set_attn_processor({n: CacheAttnProcessor2_0() for n in rp.attn_processors})
    co, cp = ou.ControlNetModel(**CFG), modeling.ControlNetModel(**CFG)
    for m, s in ((o, 0), (ro, 1), (co, 2)):
        ou.init_synthetic_(m, s)
    for m, s in ((p, 0), (rp, 1), (cp, 2)):
        modeling.init_synthetic_(m, s)
    sched = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                          clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    return (o.eval(), ro.eval(), co.eval()), (p.eval(), rp.eval(), cp.eval()), sched


def inputs(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(latents=r(1, 4, H, W), garment=r(1, 4, H, W) * 0.9, prompt=r(1, 7, 64), negative=r(1, 7, 64),
                gtok=r(1, 4, 64), pose=torch.rand(1, 3, H * 8, W * 8, generator=g))


def common(x, **kw):
    return dict(prompt=None, null_prompt=None, negative_prompt=None, ref_image=None, width=W * 8, height=H * 8,
                num_inference_steps=STEPS, output_type="latent", prompt_embeds=x["prompt"],
                negative_prompt_embeds=x["negative"], latents=x["latents"], garment_tokens=x["gtok"],
                ref_image_latents=x["garment"], **kw)


def eager(pipe):
    pipe._engine.use_cuda_graph = False
    return pipe


@torch.no_grad()
def test_base_pipeline_two_images_and_batch(emu):
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline import IMAGDressing_v1

    (o, ro, _), (p, rp, _), sched = build(emu)
    pipe = eager(IMAGDressing_v1(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None,
                                 image_encoder=None, ImgProj=None, scheduler=sched, safety_checker=None,
                                 feature_extractor=None))
    outs, refs = [], []
    for seed in (42, 43):  # two images through ONE pipeline object: no state may leak from the first to the second
        x = inputs(seed)
        ref = sample_one(o, ro, x["latents"], x["prompt"], x["negative"], x["gtok"], x["garment"], 7.5, STEPS)
        out = pipe(guidance_scale=7.5, image_scale=1.0, **common(x)).images
        assert out.shape == ref.shape and rel(out, ref) < 4e-2
        outs.append(out)
        refs.append(ref)
    xs = [inputs(42), inputs(43)]
    cat = {k: torch.cat([a[k], b[k]]) for (k, _), a, b in zip(xs[0].items(), [xs[0]] * 7, [xs[1]] * 7)}
    both = pipe(guidance_scale=7.5, image_scale=1.0, **common(cat)).images
    # (a different batch re-seeds the bf16 rounding noise, so agreement is to the bf16 floor of the 4-step chain)
    assert rel(both[0:1], refs[0]) < 4e-2 and rel(both[1:2], refs[1]) < 4e-2
    assert rel(both[0:1], outs[0]) < 4e-2 and rel(both[1:2], outs[1]) < 4e-2
    assert rel(both[0:1], outs[1]) > 0.3  # ... and the two samples are not swapped


@torch.no_grad()
def test_controlnet_and_inpainting_pipelines(emu):
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_controlnet import IMAGDressing_v1 as PControl
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_controlnet_inpainting import IMAGDressing_v1 as PInpaint

    (o, ro, co), (p, rp, cp), sched = build(emu)
    x = inputs(44)
    pipe = eager(PControl(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, controlnet=cp,
                          image_encoder=None, ImgProj=None, scheduler=sched, safety_checker=None, feature_extractor=None))
    ref = sample_one(o, ro, x["latents"], x["prompt"], x["negative"], x["gtok"], x["garment"], 7.0, STEPS, controlnet=co,
                     control_cond=x["pose"], control_scale=0.8)
    out = pipe(guidance_scale=7.0, pose_image=x["pose"], controlnet_conditioning_scale=0.8, **common(x)).images
    assert rel(out, ref) < 4e-2
    pin = eager(PInpaint(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, controlnet=cp,
                         image_encoder=None, ImgProj=None, scheduler=sched, safety_checker=None, feature_extractor=None))
    g = torch.Generator().manual_seed(49)
    img = torch.randn(1, 4, H, W, generator=g)
    mask = torch.zeros(1, 1, H, W)
    mask[..., H // 4: 3 * H // 4, W // 4: 3 * W // 4] = 1.0
    ref = sample_one(o, ro, x["latents"], x["prompt"], x["negative"], x["gtok"], x["garment"], 5.0, STEPS, controlnet=co,
                     control_cond=x["pose"], control_scale=0.5, mask=mask, image_latents=img, noise=x["latents"])
    out = pin(guidance_scale=5.0, control_image=x["pose"], strength=1.0, controlnet_conditioning_scale=0.5,
              image_latents=img, mask_latents=mask, **common(x)).images
    assert rel(out, ref) < 4e-2
    keep = (mask == 0).expand_as(out)
    assert rel(out[keep], img[keep]) < 1e-5


@torch.no_grad()
def test_ipa_controlnet_pipeline_with_face_tokens(emu):
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_ipa_controlnet import IMAGDressing_v1

    (o, ro, co), (p, rp, cp), sched = build(emu, "ipa")
    pipe = eager(IMAGDressing_v1(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, controlnet=cp,
                                 image_encoder=None, ImgProj=None, ip_ckpt=None, scheduler=sched, safety_checker=None,
                                 feature_extractor=None))
    x = inputs(46)
    gq = torch.Generator().manual_seed(47)
    face, face_null = torch.randn(1, 4, 64, generator=gq), torch.randn(1, 4, 64, generator=gq) * 0.1
    ref = sample_one(o, ro, x["latents"], torch.cat([x["prompt"], face], 1), torch.cat([x["negative"], face_null], 1),
                     x["gtok"], x["garment"], 7.0, STEPS, controlnet=co, control_cond=x["pose"], control_scale=1.0,
                     control_text=(x["prompt"], x["negative"]))
    out = pipe(guidance_scale=7.0, pose_image=x["pose"], image_scale=0.9, ipa_scale=0.8, s_lora_scale=0.2, c_lora_scale=0.3,
               face_tokens=face, face_null_tokens=face_null, **common(x)).images
    assert rel(out, ref) < 4e-2
