"""Oracle: plain-PyTorch restatement of the diffusers-0.24 SD1.5 UNet2DConditionModel / ControlNetModel that the
reference drives (reference call sites: dressing_sd/pipelines/IMAGDressing_v1_pipeline.py:466,499,511;
IMAGDressing_v1_pipeline_ipa_controlnet.py:651; inference_IMAGdressing.py:50,90). TEST INFRASTRUCTURE ONLY.

diffusers==0.24.0 (requirements.txt:12) is a third-party dependency absent from /root/reference and not
installable offline, so this file restates its published architecture with the same
state_dict key names; PARITY UNPINNED for this part (no reference test pins it). The attention arithmetic goes
through pluggable processors with the diffusers calling convention
`processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs)`, so
the reference's own adapter/attention_processor.py classes can be (and in tests are) plugged in unmodified.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

SD15 = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
            cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, sample_size=64,
            down_has_attn=(True, True, True, False), up_has_attn=(False, True, True, True), time_cond_proj_dim=None)


class Config(dict):
    __getattr__ = dict.__getitem__


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) -> [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class DefaultAttnProcessor:
    """AttnProcessor2_0 of diffusers-0.24: plain SDPA (what an un-replaced attention layer runs)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
        B, L, C = q.shape
        h = attn.heads
        q, k, v = (t.view(B, -1, h, C // h).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C)
        return attn.to_out[1](attn.to_out[0](o))


class LoRACompatibleLinear(nn.Linear):
    """diffusers-0.24 models.lora.LoRACompatibleLinear with no lora_layer attached (the reference never attaches
    one): forward(x, scale=1.0) == F.linear; RefLoraSAttnProcessor2_0 passes `scale` positionally
    (adapter/attention_processor.py:1068-1069) so the extra argument must be accepted."""

    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)


class Attention(nn.Module):
    """diffusers-0.24 models.attention_processor.Attention as configured for SD1.5: no bias on
    q/k/v, bias on to_out[0], Dropout(0), no group/spatial norm, residual_connection False, rescale 1.0."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int):
        super().__init__()
        self.heads = heads
        self.to_q = LoRACompatibleLinear(query_dim, query_dim, bias=False)
        kv = cross_attention_dim or query_dim
        self.to_k = LoRACompatibleLinear(kv, query_dim, bias=False)
        self.to_v = LoRACompatibleLinear(kv, query_dim, bias=False)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(query_dim, query_dim), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = DefaultAttnProcessor()

    def set_processor(self, processor):
        if isinstance(getattr(self, "processor", None), nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        assert attention_mask is None, "the hot path never passes a mask (SURVEY.md A.6)"
        return None

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx, cross_attention_kwargs):
        kw = cross_attention_kwargs or {}
        x = self.attn1(self.norm1(x), encoder_hidden_states=None, **kw) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=ctx, **kw) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, cross_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx, cross_attention_kwargs):
        B, C, H, W = x.shape
        r = x
        x = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            x = blk(x, ctx, cross_attention_kwargs)
        x = x.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return self.proj_out(x) + r


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.conv_shortcut(x) if self.conv_shortcut is not None else x) + h


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, groups, layers, attn, heads, cross, downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(layers)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross, groups) for _ in range(layers)])
        self.has_attn = attn
        if downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.has_down = downsample

    def forward(self, x, temb, ctx, kw):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx, kw)
            outs.append(x)
        if self.has_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb, groups, heads, cross):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cross, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])

    def forward(self, x, temb, ctx, kw):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx, kw)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb, groups, layers, attn, heads, cross, upsample):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            inp = prev if i == 0 else cout
            rs.append(ResnetBlock2D(inp + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(rs)
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross, groups) for _ in range(layers)])
        self.has_attn = attn
        if upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.has_up = upsample

    def forward(self, x, skips: List[torch.Tensor], temb, ctx, kw):
        for i, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)  # current first, skip second
            if self.has_attn:
                x = self.attentions[i](x, ctx, kw)
        if self.has_up:
            x = self.upsamplers[0](x)
        return x


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class _AttnProcessorMixin:
    @property
    def attn_processors(self) -> Dict[str, object]:
        """name -> processor in module-tree order (down, up, mid: SURVEY.md A.2)."""
        procs = {}

        def walk(name, module):
            if hasattr(module, "set_processor"):
                procs[f"{name}.processor"] = module.processor
            for sub, child in module.named_children():
                if sub == "processor":
                    continue
                walk(f"{name}.{sub}", child)

        for name, module in self.named_children():
            walk(name, module)
        return procs

    def set_attn_processor(self, processor):
        n = len(self.attn_processors)
        if isinstance(processor, dict) and len(processor) != n:
            raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                             f"not match the number of attention layers: {n}.")

        def walk(name, module):
            if hasattr(module, "set_processor"):
                module.set_processor(processor.pop(f"{name}.processor") if isinstance(processor, dict) else processor)
            for sub, child in module.named_children():
                if sub == "processor":
                    continue
                walk(f"{name}.{sub}", child)

        processor = dict(processor) if isinstance(processor, dict) else processor
        for name, module in self.named_children():
            walk(name, module)


class UNet2DConditionModel(nn.Module, _AttnProcessorMixin):
    def __init__(self, **overrides):
        super().__init__()
        cfg = Config({**SD15, **overrides})
        self.config = cfg
        boc = cfg.block_out_channels
        temb = boc[0] * 4
        g, heads, cross, layers = cfg.norm_num_groups, cfg.attention_head_dim, cfg.cross_attention_dim, cfg.layers_per_block
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        # ModuleLists for down/up are registered before mid_block -> attn_processors order down, up, mid
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        out = boc[0]
        for i, c in enumerate(boc):
            inp, out = out, c
            self.down_blocks.append(DownBlock(inp, out, temb, g, layers, cfg.down_has_attn[i], heads, cross, i < len(boc) - 1))
        self.mid_block = MidBlock(boc[-1], temb, g, heads, cross)
        rev = list(reversed(boc))
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            inp = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(inp, out, prev, temb, g, layers + 1, cfg.up_has_attn[i], heads, cross, i < len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    @property
    def in_channels(self):
        return self.config.in_channels

    def time_embed(self, sample, timestep):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=sample.device)
        t = t.reshape(-1).to(sample.device).expand(sample.shape[0]) if t.numel() == 1 else t.to(sample.device)
        emb = timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype)
        return self.time_embedding(emb)

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None, **unused):
        temb = self.time_embed(sample, timestep)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states, cross_attention_kwargs)
            skips += outs
        if down_block_additional_residuals is not None:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        x = self.mid_block(x, temb, encoder_hidden_states, cross_attention_kwargs)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states, cross_attention_kwargs)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x,)


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, out_ch, cond_ch=3, chans=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_ch, chans[0], 3, padding=1)
        self.blocks = nn.ModuleList([])
        for i in range(len(chans) - 1):
            self.blocks.append(nn.Conv2d(chans[i], chans[i], 3, padding=1))
            self.blocks.append(nn.Conv2d(chans[i], chans[i + 1], 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(chans[-1], out_ch, 3, padding=1)

    def forward(self, c):
        x = F.silu(self.conv_in(c))
        for b in self.blocks:
            x = F.silu(b(x))
        return self.conv_out(x)


class ControlNetModel(nn.Module, _AttnProcessorMixin):
    """diffusers-0.24 ControlNetModel (v1.1 SD1.5): conditioning embedding + UNet encoder half + mid + 13 zero convs."""

    def __init__(self, **overrides):
        super().__init__()
        cfg = Config({**SD15, "global_pool_conditions": False, **overrides})
        self.config = cfg
        boc = cfg.block_out_channels
        temb = boc[0] * 4
        g, heads, cross, layers = cfg.norm_num_groups, cfg.attention_head_dim, cfg.cross_attention_dim, cfg.layers_per_block
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(boc[0])
        self.down_blocks = nn.ModuleList([])
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(boc[0], boc[0], 1)])
        out = boc[0]
        for i, c in enumerate(boc):
            inp, out = out, c
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(inp, out, temb, g, layers, cfg.down_has_attn[i], heads, cross, not last))
            for _ in range(layers + (0 if last else 1)):
                self.controlnet_down_blocks.append(nn.Conv2d(out, out, 1))
        self.controlnet_mid_block = nn.Conv2d(boc[-1], boc[-1], 1)
        self.mid_block = MidBlock(boc[-1], temb, g, heads, cross)

    time_embed = UNet2DConditionModel.time_embed

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0,
                guess_mode=False, return_dict=False, **unused):
        temb = self.time_embed(sample, timestep)
        x = self.conv_in(sample) + self.controlnet_cond_embedding(controlnet_cond)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states, None)
            skips += outs
        x = self.mid_block(x, temb, encoder_hidden_states, None)
        down = [conv(s) * conditioning_scale for s, conv in zip(skips, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(x) * conditioning_scale
        return down, mid


def init_synthetic_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Deterministic synthetic weights (no checkpoints offline): each parameter is drawn from a
    generator seeded by (seed, crc32(parameter name)), so any two module trees with the same state_dict keys get
    identical values regardless of registration order. Fan-in scaled N(0, 1/fan_in) for matrices / convs, small
    biases, norm gamma ~ 1 — keeps activations O(1) through the 60-odd layers."""
    import math
    import zlib

    with torch.no_grad():
        for name, p in model.named_parameters():
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
            if p.dim() >= 2:
                v = torch.randn(p.shape, generator=g) * (1.0 / math.sqrt(p[0].numel()))
            elif "norm" in name and name.endswith("weight"):
                v = 1.0 + 0.05 * torch.randn(p.shape, generator=g)
            else:
                v = 0.02 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.dtype))
    if hasattr(model, "invalidate_packed"):
        model.invalidate_packed()
    return model
