"""CPU: the reference-facing import surface: module paths, class names, constructor / call signatures,
state_dict keys and isinstance behaviour of the drop-in modules, plus the opt-in `diffusers` stand-in."""
import inspect
import json
import os

import pytest
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_module_paths_and_signatures():
    import adapter.attention_processor as ap
    from adapter.resampler import ProjPlusModel, Resampler  # noqa: F401
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline import IMAGDressing_v1 as P0
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_controlnet import IMAGDressing_v1 as P1
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_controlnet_inpainting import IMAGDressing_v1 as P3
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline_ipa_controlnet import IMAGDressing_v1 as P2

    for P in (P0, P1, P2, P3):
        assert P.__name__ == "IMAGDressing_v1"
    # constructor kwargs the scripts pass (inference_IMAGdressing.py:129-134, ..._ipa_controlnetpose.py:142-151)
    assert list(inspect.signature(P0.__init__).parameters)[1:11] == [
        "vae", "reference_unet", "unet", "tokenizer", "text_encoder", "image_encoder", "ImgProj", "scheduler",
        "safety_checker", "feature_extractor"]
    assert "controlnet" in inspect.signature(P1.__init__).parameters
    assert "ip_ckpt" in inspect.signature(P2.__init__).parameters
    call = inspect.signature(P0.__call__).parameters
    for k in ("prompt", "null_prompt", "negative_prompt", "ref_image", "width", "height", "num_inference_steps",
              "guidance_scale", "ref_clip_image", "image_scale", "generator", "prompt_embeds", "negative_prompt_embeds"):
        assert k in call
    for k in ("pose_image", "face_clip_image", "faceid_embeds", "ipa_scale", "s_lora_scale", "c_lora_scale",
              "controlnet_conditioning_scale"):
        assert k in inspect.signature(P2.__call__).parameters
    for k in ("image", "mask_image", "control_image", "strength"):
        assert k in inspect.signature(P3.__call__).parameters
    # processors: names, ctor args, parameter names, class identity used by set_scale / set_ipa_scale
    refs = ap.RefSAttnProcessor2_0("n", 320, scale=0.5)
    assert list(refs.state_dict()) == ["to_k_ref.weight", "to_v_ref.weight"] and refs.scale == 0.5 and refs.name == "n"
    ip = ap.LoRAIPAttnProcessor2_0(320, 768, rank=128, num_tokens=4)
    assert {"to_k_ip.weight", "to_v_ip.weight", "to_q_lora.down.weight", "to_q_lora.up.weight", "to_out_lora.up.weight"} \
        <= set(ip.state_dict())
    assert ip.state_dict()["to_q_lora.down.weight"].shape == (128, 320)
    lr = ap.LoraRefSAttnProcessor2_0("n", 640)
    assert lr.rank == 128 and not isinstance(ap.RefLoraSAttnProcessor2_0("n", 640), ap.LoraRefSAttnProcessor2_0)
    assert isinstance(ap.CacheAttnProcessor2_0().cache, dict)


def test_unet_host_surface():
    from adapter.attention_processor import CacheAttnProcessor2_0
    from imagdressing_b200.modeling import ControlNetModel, UNet2DConditionModel

    cfg = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=8)
    u = UNet2DConditionModel(**cfg)
    names = list(u.attn_processors)
    assert len(names) == 32 and names[0] == "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor"
    assert names[-1] == "mid_block.attentions.0.transformer_blocks.0.attn2.processor"  # down, up, mid order
    assert u.config.cross_attention_dim == 64 and u.in_channels == 4
    u.set_attn_processor({n: CacheAttnProcessor2_0() for n in names})
    assert all(isinstance(p, CacheAttnProcessor2_0) for p in u.attn_processors.values())
    try:
        u.set_attn_processor({names[0]: CacheAttnProcessor2_0()})
        raise AssertionError("a short processor dict must be rejected")
    except ValueError:
        pass
    sd = u.state_dict()
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.time_emb_proj.bias",
              "down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight",
              "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.weight", "mid_block.resnets.1.conv2.weight",
              "conv_norm_out.weight", "conv_out.bias"):
        assert k in sd, k
    u2 = UNet2DConditionModel(**cfg)
    u2.load_state_dict(sd)
    c = ControlNetModel(**cfg)
    assert len(c.controlnet_down_blocks) == 12 and "controlnet_cond_embedding.conv_in.weight" in c.state_dict()
    assert c.config.global_pool_conditions is False
    # processor modules become sub-modules (so .to() / state_dict() include them), as diffusers' set_processor does
    from adapter.attention_processor import RefSAttnProcessor2_0

    u.set_attn_processor({n: (RefSAttnProcessor2_0(n, 32) if "attn1" in n else CacheAttnProcessor2_0()) for n in names})
    assert any(k.endswith("attn1.processor.to_k_ref.weight") for k in u.state_dict())


def test_diffusers_stand_in_is_opt_in(tmp_path):
    """`import diffusers` must NOT resolve to the stand-in unless imagdressing_b200/compat is put on the path."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from diffusers import UNet2DConditionModel, DDIMScheduler, ControlNetModel, AutoencoderKL;"
            "from diffusers.utils import load_image, is_accelerate_available;"
            "from diffusers.pipelines.stable_diffusion import StableDiffusionSafetyChecker;"
            "import imagdressing_b200.modeling as m; assert UNet2DConditionModel is m.UNet2DConditionModel;"
            "s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear',"
            " clip_sample=False, set_alpha_to_one=False, steps_offset=1); s.set_timesteps(50);"
            "assert s.timesteps[0] == 981; print('ok')") % (ROOT, os.path.join(ROOT, "imagdressing_b200", "compat"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
    cfg_dir = tmp_path / "unet"
    cfg_dir.mkdir()
    json.dump({"block_out_channels": [32, 64, 64, 64], "cross_attention_dim": 64, "norm_num_groups": 8},
              open(cfg_dir / "config.json", "w"))
    from imagdressing_b200.modeling import UNet2DConditionModel

    with pytest.raises(FileNotFoundError):  # a directory without weights must not silently yield a random UNet
        UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    u = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet", torch_dtype=torch.float16,
                                             allow_random_init=True)
    assert u.dtype == torch.float16 and u.config.block_out_channels == (32, 64, 64, 64)
    # weights round-trip through both supported file formats
    from safetensors.torch import save_file

    sd = {k: v.float().contiguous() for k, v in u.state_dict().items()}
    save_file(sd, str(cfg_dir / "diffusion_pytorch_model.safetensors"))
    u2 = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert all(torch.equal(u2.state_dict()[k], v) for k, v in sd.items())
    os.remove(cfg_dir / "diffusion_pytorch_model.safetensors")
    torch.save(sd, str(cfg_dir / "diffusion_pytorch_model.bin"))
    u3 = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert all(torch.equal(u3.state_dict()[k], v) for k, v in sd.items())
