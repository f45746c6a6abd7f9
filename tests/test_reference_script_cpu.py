"""CPU, BASELINE.json configs[0] / VERDICT r1 item 7-8: the reference's UNMODIFIED `inference_IMAGdressing.py` is executed
as `__main__` (runpy) — `prepare()` + `pipe(...)` + the image grid it saves — against this repo's drop-ins:

  * `dressing_sd.pipelines.IMAGDressing_v1_pipeline.IMAGDressing_v1`, `adapter.attention_processor.*`,
    `adapter.resampler.Resampler` resolve to this repo (same module paths);
  * `from diffusers import UNet2DConditionModel, AutoencoderKL, DDIMScheduler` resolves to the opt-in stand-in
    (imagdressing_b200/compat), i.e. the kernel-backed UNet / VAE / scheduler;
  * `transformers` (real, installed) provides CLIPTextModel / CLIPTokenizer / CLIPVisionModelWithProjection /
    CLIPImageProcessor; torchvision (real) the transforms.

Test infrastructure only: the hub ids the script hard-codes ("stabilityai/sd-vae-ft-mse", "SG161222/Realistic_Vision_V4.0_noVAE",
"h94/IP-Adapter") are satisfied by directories of those names under a temporary working directory holding small random-init
models (a relative path that exists is a local directory for both transformers and our from_pretrained), a toy CLIP BPE
vocabulary, and a synthetic `{"module": ...}` checkpoint with the `ref_unet.` / `unet.` / `proj.` / `adapter_modules.`
layout; the kernel wrappers are emulated in torch (tests/emulated_ops.py) because this suite has no GPU. The script itself
is read from /root/reference at run time and is not copied; the test is skipped where the reference tree does not exist
(the GPU box).
"""
import json
import os
import runpy
import sys

import pytest
import torch

import emulated_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = "/root/reference/inference_IMAGdressing.py"
pytestmark = pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference tree not present")

UNET_CFG = dict(block_out_channels=[32, 64, 64, 64], cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
VAE_CFG = dict(block_out_channels=[32, 32, 64, 64], norm_num_groups=8)


def make_tokenizer(d):
    """A CLIP-style byte-level BPE vocabulary with NO merges (every byte, bare and word-final), built and saved through the
    installed transformers so the files are in whatever format that version reads back."""
    from transformers import CLIPTokenizer

    # GPT-2 / CLIP byte -> printable unicode table
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    chars = [chr(c) for c in cs]
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    tok = CLIPTokenizer(vocab=vocab, merges=[], model_max_length=77)
    tok.save_pretrained(d)
    ids = CLIPTokenizer.from_pretrained(d)("a cat", padding="max_length", max_length=77, truncation=True).input_ids
    assert len(ids) == 77 and ids[0] == vocab["<|startoftext|>"] and vocab["<|endoftext|>"] in ids
    return len(vocab)


def make_hub(tmp):
    from safetensors.torch import save_file
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModelWithProjection

    from imagdressing_b200 import modeling, vae

    torch.manual_seed(0)
    rv = os.path.join(tmp, "SG161222", "Realistic_Vision_V4.0_noVAE")
    n_vocab = make_tokenizer(os.path.join(rv, "tokenizer"))
    CLIPTextModel(CLIPTextConfig(vocab_size=n_vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                 num_attention_heads=4, max_position_embeddings=77, projection_dim=64,
                                 bos_token_id=n_vocab - 2, eos_token_id=n_vocab - 1, pad_token_id=n_vocab - 1)
                  ).save_pretrained(os.path.join(rv, "text_encoder"))
    unet = modeling.UNet2DConditionModel(**UNET_CFG)
    modeling.init_synthetic_(unet, 0)
    os.makedirs(os.path.join(rv, "unet"))
    json.dump(UNET_CFG, open(os.path.join(rv, "unet", "config.json"), "w"))
    save_file({k: v.contiguous() for k, v in unet.state_dict().items()},
              os.path.join(rv, "unet", "diffusion_pytorch_model.safetensors"))
    CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=48, intermediate_size=96, num_hidden_layers=2,
                                                   num_attention_heads=4, image_size=224, patch_size=14, projection_dim=32)
                                  ).save_pretrained(os.path.join(tmp, "h94", "IP-Adapter", "models", "image_encoder"))
    v = vae.AutoencoderKL(**VAE_CFG)
    modeling.init_synthetic_(v, 3)
    vd = os.path.join(tmp, "stabilityai", "sd-vae-ft-mse")
    os.makedirs(vd)
    json.dump(VAE_CFG, open(os.path.join(vd, "config.json"), "w"))
    save_file({k: v_.contiguous() for k, v_ in v.state_dict().items()}, os.path.join(vd, "diffusion_pytorch_model.safetensors"))
    return unet


def make_checkpoint(tmp, unet):
    """SDModel-style DeepSpeed checkpoint (train.py:244-253 member names) with fresh 'trained' values."""
    from adapter.attention_processor import CAttnProcessor2_0, RefSAttnProcessor2_0
    from adapter.resampler import Resampler
    from imagdressing_b200 import modeling

    g = torch.Generator().manual_seed(1)
    ref = modeling.UNet2DConditionModel(**UNET_CFG)
    modeling.init_synthetic_(ref, 1)
    proj = Resampler(dim=64, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=48, output_dim=64, ff_mult=4)
    sd = {"ref_unet." + k: v for k, v in ref.state_dict().items()}
    sd.update({"unet." + k: v for k, v in unet.state_dict().items()})
    sd.update({"proj." + k: v + 0.01 * torch.randn(v.shape, generator=g) for k, v in proj.state_dict().items()})
    boc = UNET_CFG["block_out_channels"]
    for i, name in enumerate(unet.attn_processors.keys()):
        if name.endswith("attn1.processor"):
            hidden = boc[-1] if name.startswith("mid") else (list(reversed(boc))[int(name[len("up_blocks.")])]
                                                             if name.startswith("up") else boc[int(name[len("down_blocks.")])])
            for w in ("to_k_ref.weight", "to_v_ref.weight"):
                sd[f"adapter_modules.{i}.{w}"] = torch.randn(hidden, hidden, generator=g) * hidden ** -0.5
    path = os.path.join(tmp, "ckpt", "IMAGDressing-v1_512.pt")
    os.makedirs(os.path.dirname(path))
    torch.save({"module": sd}, path)
    return path


def test_unmodified_inference_script_runs_end_to_end(tmp_path, monkeypatch):
    import numpy as np
    from PIL import Image

    tmp = str(tmp_path)
    emulated_ops.install(monkeypatch)
    unet = make_hub(tmp)
    ckpt = make_checkpoint(tmp, unet)
    cloth = os.path.join(tmp, "cloth.png")
    rng = np.random.default_rng(0)
    Image.fromarray(rng.integers(0, 255, (300, 240, 3), dtype=np.uint8)).save(cloth)

    # the stand-in diffusers is opt-in: visible only inside this test
    compat = os.path.join(ROOT, "imagdressing_b200", "compat")
    monkeypatch.syspath_prepend(compat)
    monkeypatch.syspath_prepend(ROOT)
    for m in [m for m in sys.modules if m == "diffusers" or m.startswith("diffusers.")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.chdir(tmp)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    monkeypatch.setattr(sys, "argv", ["inference_IMAGdressing.py", "--cloth_path", cloth, "--model_ckpt", ckpt,
                                      "--output_path", os.path.join(tmp, "out"), "--device", "cpu"])
    seen = {}
    import imagdressing_b200.pipelines as pl

    real_run = pl._DressingPipelineBase._run

    def spy(self, **kw):  # observe (not alter) what the script asked for
        seen.update(steps=kw["num_inference_steps"], size=(kw["height"], kw["width"]), guidance=kw["guidance_scale"],
                    prompt=kw["prompt"], has_clip=kw["ref_clip_image"] is not None)
        seen["garment_unet_loaded"] = float(self.reference_unet.conv_in.weight.float().abs().sum())
        return real_run(self, **kw)

    monkeypatch.setattr(pl._DressingPipelineBase, "_run", spy)
    try:
        runpy.run_path(SCRIPT, run_name="__main__")
    finally:
        for m in [m for m in sys.modules if m == "diffusers" or m.startswith("diffusers.")]:
            sys.modules.pop(m, None)
    out = os.path.join(tmp, "out", "cloth.png")
    assert os.path.exists(out)
    grid = Image.open(out)
    assert grid.size == (1024, 640)  # [resized garment | generated 512x640 image]
    arr = np.asarray(grid)[:, 512:]
    assert arr.std() > 1.0  # a real decoded image, not a constant
    assert seen == dict(steps=50, size=(640, 512), guidance=7.5, prompt="A beautiful woman, best quality, high quality",
                        has_clip=True, garment_unet_loaded=seen["garment_unet_loaded"])
    # the garment UNet really carries the checkpoint's `ref_unet.` weights (seed 1), not the base UNet's (seed 0)
    from imagdressing_b200 import modeling

    ref = modeling.UNet2DConditionModel(**UNET_CFG)
    modeling.init_synthetic_(ref, 1)
    want = float(ref.conv_in.weight.half().float().abs().sum())
    assert abs(seen["garment_unet_loaded"] - want) < 1e-3 * want
