"""CPU: the kernel-backed CLIP text / vision encoders (imagdressing_b200/clip.py) with the kernel
wrappers emulated in torch against the PINNED oracle for this row — the installed `transformers` implementation itself
(CLIPTextModel / CLIPVisionModelWithProjection, the classes inference_IMAGdressing.py:44-49 constructs): token + position
embedding, causal attention, quick_gelu / gelu MLPs, patch embedding as a GEMM, class token, pre-LN, `hidden_states[-2]`."""
import pytest
import torch

import emulated_ops


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture
def emu(monkeypatch):
    emulated_ops.install(monkeypatch)
    from imagdressing_b200 import clip

    return clip


@torch.no_grad()
def test_text_encoder_matches_transformers(emu):
    from transformers import CLIPTextConfig, CLIPTextModel

    torch.manual_seed(0)
    hf = CLIPTextModel(CLIPTextConfig(vocab_size=300, hidden_size=128, intermediate_size=256, num_hidden_layers=3,
                                      num_attention_heads=2, max_position_embeddings=77, hidden_act="quick_gelu",
                                      eos_token_id=299, bos_token_id=298, pad_token_id=299)).eval()
    enc = emu.accelerate(hf)
    assert isinstance(enc, emu.ClipTextEncoder) and emu.accelerate(enc) is enc and enc.config is hf.config
    ids = torch.randint(0, 298, (2, 77))
    ids[:, 0], ids[:, 20:] = 298, 299
    want = hf(ids, output_hidden_states=True)
    got = enc(ids, output_hidden_states=True)
    assert rel(got[0], want[0]) < 2e-2 and rel(enc(ids)[0], want.last_hidden_state) < 2e-2
    assert len(got.hidden_states) == len(want.hidden_states) == 4
    for a, b in zip(got.hidden_states, want.hidden_states):
        assert rel(a, b) < 2e-2
    # causality: changing a late token must not change earlier positions
    ids2 = ids.clone()
    ids2[:, 15] = 7
    out2 = enc(ids2)[0]
    assert torch.equal(out2[:, :15], got[0][:, :15]) and rel(out2[:, 15:], got[0][:, 15:]) > 1e-3
    # in-place weight edits are picked up (packs are versioned on the parameters)
    hf.text_model.final_layer_norm.weight.mul_(2.0)
    assert rel(enc(ids)[0], hf(ids)[0]) < 2e-2


@torch.no_grad()
def test_vision_encoder_matches_transformers(emu):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    torch.manual_seed(1)
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=160, intermediate_size=320, num_hidden_layers=4,
                                                        num_attention_heads=2, image_size=56, patch_size=14,
                                                        projection_dim=64, hidden_act="gelu")).eval()
    enc = emu.accelerate(hf)
    assert isinstance(enc, emu.ClipVisionEncoder) and enc.config.hidden_size == 160
    px = torch.randn(2, 3, 56, 56)
    want = hf(px, output_hidden_states=True).hidden_states
    got = enc(px, output_hidden_states=True).hidden_states
    assert len(got) == len(want) == 5 and got[-1] is None  # the last layer is never needed for hidden_states[-2]
    assert rel(got[-2], want[-2]) < 2e-2 and rel(got[0], want[0]) < 2e-2 and got[-2].shape == (2, 17, 160)
    assert rel(enc(px).last_hidden_state, hf.vision_model.encoder(want[0]).last_hidden_state) < 3e-2


@torch.no_grad()
def test_pipeline_uses_the_encoders(emu, monkeypatch):
    """encode_prompt / the garment-token branch of the pipeline (IMAGDressing_v1_pipeline.py:396-415) through the wrappers."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    from adapter.resampler import Resampler
    from dressing_sd.pipelines.IMAGDressing_v1_pipeline import IMAGDressing_v1
    from imagdressing_b200 import clip, modeling
    from test_pipelines_cpu import build

    (_, _, _), (p, rp, _), sched = build(modeling)
    torch.manual_seed(2)
    vis = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=160, intermediate_size=320, num_hidden_layers=3,
                                                         num_attention_heads=2, image_size=56, patch_size=14,
                                                         projection_dim=64)).eval()
    proj = Resampler(dim=64, depth=2, dim_head=16, heads=4, num_queries=4, embedding_dim=160, output_dim=64, ff_mult=2)
    pipe = IMAGDressing_v1(vae=None, reference_unet=rp, unet=p, tokenizer=None, text_encoder=None,
                           image_encoder=clip.accelerate(vis), ImgProj=proj, scheduler=sched, safety_checker=None,
                           feature_extractor=None)
    px = torch.randn(1, 3, 56, 56)
    tok = pipe._garment_tokens(px, None, torch.device("cpu"), torch.float32)
    want = proj(vis(px, output_hidden_states=True).hidden_states[-2])
    assert tok.shape == (1, 4, 64) and rel(tok, want) < 3e-2
