"""GPU parity of the kernel-backed CLIP encoders at the REAL sizes the reference loads (inference_IMAGdressing.py:44-49):
the SD1.5 text encoder (CLIP ViT-L/14 text tower: 12 layers, width 768, 12 heads of 64, quick_gelu, 77 tokens, causal) and
the IP-Adapter image encoder (CLIP ViT-H/14: 32 layers, width 1280, 16 heads of 80, gelu, 257 tokens), random-init.
Oracle = the installed transformers implementation in fp32 on the same GPU — a PINNED oracle (third-party code, not a
restatement); tolerance calibrated against the same module run in torch bf16."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@torch.no_grad()
def test_text_encoder_vit_l(cuda_device):
    from transformers import CLIPTextConfig, CLIPTextModel

    from imagdressing_b200 import clip

    torch.manual_seed(0)
    hf = CLIPTextModel(CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
                       ).to(cuda_device).eval()
    ids = torch.randint(0, 49406, (4, 77), device=cuda_device)
    ids[:, 0], ids[:, 30:] = 49406, 49407
    want = hf(ids, output_hidden_states=True)
    enc = clip.auto_accelerate(hf)
    assert isinstance(enc, clip.ClipTextEncoder)
    got = enc(ids, output_hidden_states=True)
    drift = rel_l2(hf.bfloat16()(ids)[0], want[0])
    hf.float()
    e = rel_l2(got[0], want[0])
    e2 = rel_l2(got.hidden_states[-2], want.hidden_states[-2])
    print(f"CLIP text (ViT-L/14 tower): last_hidden_state rel-L2 {e:.4f} (torch-bf16 {drift:.4f}), hidden_states[-2] {e2:.4f}")
    assert e < max(2 * drift, 2e-2) and e2 < max(2 * drift, 2e-2)
    assert torch.equal(enc(ids)[0], got[0])


@torch.no_grad()
def test_vision_encoder_vit_h(cuda_device):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    from imagdressing_b200 import clip

    torch.manual_seed(1)
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                                                        num_attention_heads=16, image_size=224, patch_size=14,
                                                        projection_dim=1024, hidden_act="gelu")).to(cuda_device).eval()
    px = torch.randn(2, 3, 224, 224, device=cuda_device)
    want = hf(px, output_hidden_states=True).hidden_states[-2]
    drift = rel_l2(hf.bfloat16()(px.bfloat16(), output_hidden_states=True).hidden_states[-2], want)
    hf.float()
    enc = clip.auto_accelerate(hf)
    got = enc(px, output_hidden_states=True).hidden_states[-2]
    e = rel_l2(got, want)
    print(f"CLIP vision (ViT-H/14): hidden_states[-2] [2,257,1280] rel-L2 {e:.4f} (torch-bf16 {drift:.4f})")
    assert got.shape == (2, 257, 1280) and e < max(2 * drift, 2e-2)


def test_causal_attention_kernel(cuda_device):
    """Causal mask of the attention kernel on its own (77 tokens in one key block; 200 tokens over two blocks)."""
    from imagdressing_b200 import ops

    for L, heads, hd in ((77, 12, 64), (200, 4, 80), (128, 2, 64)):
        C = heads * hd
        g = torch.Generator().manual_seed(L)
        qkv = torch.randn(2, L, 3 * C, generator=g).to(cuda_device).bfloat16()
        flat = lambda t: t.as_strided((t.shape[0] * t.shape[1], t.shape[2]), (t.stride(1), 1), t.storage_offset())
        s0 = ops.kv_stream(flat(qkv[..., C:2 * C]), flat(qkv[..., 2 * C:]), L)
        out = ops.attention(flat(qkv[..., :C]), 2, L, heads, hd, s0, causal=True).view(2, L, C)
        q, k, v = (t.float().view(2, L, heads, hd).transpose(1, 2) for t in (qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]))
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(2, L, C)
        assert rel_l2(out, ref) < 1e-2, (L, heads, hd, rel_l2(out, ref))
