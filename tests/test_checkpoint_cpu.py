"""CPU: the checkpoint-format loader (imagdressing_b200/checkpoint.py) — a synthetic DeepSpeed
`{"module": ...}` checkpoint with the `ref_unet.` / `unet.` / `proj.` / `adapter_modules.` layout routed as
inference_IMAGdressing.py:97-114 routes it (incl. quirk B8: `unet.*` ignored), a FaceID `.bin` / `.safetensors` checkpoint
routed as IMAGDressing_v1_pipeline_ipa_controlnet.py:88-101, and — the point of the row — the packed / LoRA-merged weights
the KERNELS see afterwards: the pipeline (kernel wrappers emulated) must match the oracle loaded with the same files,
also when the load happens AFTER a forward has already packed the old weights."""
import pytest
import torch
import torch.nn as nn

import emulated_ops
from oracle import processors as op
from oracle.pipeline import sample_one
from test_pipelines_cpu import CFG, STEPS, build, common, eager, inputs, rel


@pytest.fixture
def emu(monkeypatch):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import modeling

    return modeling


def resamplers():
    from adapter.resampler import Resampler

    kw = dict(dim=64, depth=2, dim_head=16, heads=4, num_queries=4, embedding_dim=48, output_dim=64, ff_mult=2)
    return Resampler(**kw), op.Resampler(**kw)


@torch.no_grad()
def test_module_checkpoint_routing_and_packed_weights(emu, tmp_path):
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline import IMAGDressing_v1
    from imagdressing_b200 import checkpoint as ck

    (o, ro, _), (p, rp, _), sched = build(emu)
    proj_p, proj_o = resamplers()
    pipe = eager(IMAGDressing_v1(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, image_encoder=None,
                                 ImgProj=proj_p, scheduler=sched, safety_checker=None, feature_extractor=None))
    x = inputs(70)
    before = pipe(guidance_scale=7.5, **common(x)).images  # packs the ORIGINAL weights

    # a "trained" checkpoint: new garment UNet, new adapters, new resampler, and a DIFFERENT denoising UNet that the
    # reference never loads (B8)
    g = torch.Generator().manual_seed(1)
    rnd = lambda sd, s=0.05: {k: (v + s * torch.randn(v.shape, generator=g)).to(v.dtype) for k, v in sd.items()}
    new_ref = rnd(rp.state_dict())
    new_proj = rnd(proj_p.state_dict())
    new_adapters = rnd(ck.processor_list(p).state_dict(), 0.1)
    assert any(k.endswith("to_k_ref.weight") for k in new_adapters) and "0.to_k_ref.weight" in new_adapters
    assert not any(k.startswith("1.") for k in new_adapters)  # attn2 (CAttn) processors hold no parameters
    decoy_unet = rnd({k: v for k, v in p.state_dict().items() if ".processor." not in k}, 1.0)
    module = {**{"ref_unet." + k: v for k, v in new_ref.items()}, **{"unet." + k: v for k, v in decoy_unet.items()},
              **{"proj." + k: v for k, v in new_proj.items()}, **{"adapter_modules." + k: v for k, v in new_adapters.items()},
              "stray.key": torch.zeros(1)}
    path = tmp_path / "IMAGDressing-v1_512.pt"
    torch.save({"module": module}, path)

    parts, other = ck.split_module_checkpoint(module)
    assert other == ["stray.key"] and set(parts) == {"ref_unet", "unet", "proj", "adapter_modules"}
    unet_before = {k: v.clone() for k, v in p.state_dict().items()}
    rep = ck.load_module_checkpoint(str(path), reference_unet=rp, image_proj=proj_p, unet=p)
    assert rep["unet_ignored"] == len(decoy_unet) and rep["unrouted"] == ["stray.key"]
    for k, v in p.state_dict().items():  # B8: the denoising UNet's own weights are untouched ...
        if ".processor." not in k:
            assert torch.equal(v, unet_before[k])
    assert all(torch.equal(rp.state_dict()[k], v) for k, v in new_ref.items())
    assert all(torch.equal(proj_p.state_dict()[k], v) for k, v in new_proj.items())
    # ... the oracle side, loaded the way the reference script does it
    ro.load_state_dict(new_ref)
    proj_o.load_state_dict(new_proj)
    nn.ModuleList(list(o.attn_processors.values())).load_state_dict(new_adapters)
    ref = sample_one(o, ro, x["latents"], x["prompt"], x["negative"], x["gtok"], x["garment"], 7.5, STEPS)
    after = pipe(guidance_scale=7.5, **common(x)).images
    assert rel(after, ref) < 4e-2 and rel(before, ref) > 2 * rel(after, ref)  # no stale packed weights
    tok = torch.randn(2, 9, 48, generator=g)
    assert rel(proj_p(tok), proj_o(tok)) < 2e-2

    # round trip through the exporter
    out = tmp_path / "export.pt"
    ck.export_module_checkpoint(str(out), reference_unet=rp, image_proj=proj_p, unet=p)
    sd2 = torch.load(out, weights_only=True)["module"]
    assert all(torch.equal(sd2["ref_unet." + k], v) for k, v in new_ref.items())
    assert all(torch.equal(sd2["adapter_modules." + k], v) for k, v in new_adapters.items())
    assert "unet.conv_in.weight" in sd2 and "proj.latents" in sd2
    # train.py:355-359 initialisation helper
    assert ck.copy_self_attention_into_ref_projections(p) == 16
    name = "down_blocks.0.attentions.0.transformer_blocks.0.attn1"
    assert torch.equal(p.attn_processors[name + ".processor"].to_k_ref.weight, p.state_dict()[name + ".to_k.weight"])


@pytest.mark.parametrize("fmt", ["bin", "safetensors"])
@torch.no_grad()
def test_faceid_checkpoint_formats(emu, tmp_path, fmt):
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_ipa_controlnet import IMAGDressing_v1
    from imagdressing_b200 import checkpoint as ck

    (o, ro, co), (p, rp, cp), sched = build(emu, "ipa")
    pipe = eager(IMAGDressing_v1(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None, controlnet=cp,
                                 image_encoder=None, ImgProj=None, ip_ckpt=None, scheduler=sched, safety_checker=None,
                                 feature_extractor=None))
    g = torch.Generator().manual_seed(2)
    layers = ck.processor_list(p)
    ip = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in layers.state_dict().items()
          if "to_k_ip" in k or "to_v_ip" in k or "_lora" in k}
    assert ip and all(int(k.split(".")[0]) % 2 == 1 or "_lora" in k for k in ip)  # to_k_ip / to_v_ip live at ODD indices
    proj = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in pipe.image_proj_model.state_dict().items()}
    path = tmp_path / f"ip-adapter-faceid-plusv2_sd15.{fmt}"
    if fmt == "bin":
        torch.save({"image_proj": proj, "ip_adapter": ip}, path)
    else:
        from safetensors.torch import save_file

        save_file({**{"image_proj." + k: v for k, v in proj.items()}, **{"ip_adapter." + k: v for k, v in ip.items()}},
                  str(path))
    pipe.ip_ckpt = str(path)
    pipe.load_ip_adapter()
    got = ck.processor_list(p).state_dict()
    assert all(torch.equal(got[k], v) for k, v in ip.items())
    assert all(torch.equal(pipe.image_proj_model.state_dict()[k], v) for k, v in proj.items())
    nn.ModuleList(list(o.attn_processors.values())).load_state_dict(ip, strict=False)
    x = inputs(71)
    gq = torch.Generator().manual_seed(3)
    face, face_null = torch.randn(1, 4, 64, generator=gq), torch.randn(1, 4, 64, generator=gq) * 0.1
    ref = sample_one(o, ro, x["latents"], torch.cat([x["prompt"], face], 1), torch.cat([x["negative"], face_null], 1),
                     x["gtok"], x["garment"], 7.0, STEPS, controlnet=co, control_cond=x["pose"], control_scale=1.0,
                     control_text=(x["prompt"], x["negative"]))
    out = pipe(guidance_scale=7.0, pose_image=x["pose"], image_scale=0.9, ipa_scale=0.8, s_lora_scale=0.2, c_lora_scale=0.3,
               face_tokens=face, face_null_tokens=face_null, **common(x)).images
    assert rel(out, ref) < 4e-2
