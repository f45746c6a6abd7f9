"""Oracle: plain-PyTorch restatement of the diffusers-0.24 `AutoencoderKL` the reference pipelines call at their edges —
`vae.encode(ref_image).latent_dist.mean * 0.18215` (dressing_sd/pipelines/IMAGDressing_v1_pipeline.py:454-458),
`vae.decode(latents / scaling_factor)` (:544), `latent_dist.sample(generator)` in the inpainting pipeline
(IMAGDressing_v1_pipeline_controlnet_inpainting.py `prepare_latents` -> `_encode_vae_image`). The checkpoint the scripts
name is `stabilityai/sd-vae-ft-mse` (inference_IMAGdressing.py:42): block_out_channels (128, 256, 512, 512),
layers_per_block 2, latent_channels 4, norm_num_groups 32, SiLU, scaling_factor 0.18215. TEST INFRASTRUCTURE ONLY.

diffusers is absent offline: restated from the published architecture with the diffusers state_dict
key names — PARITY UNPINNED for this file (no reference test pins it):
  encoder: conv_in 3->128; 4 x DownEncoderBlock2D (2 ResnetBlock2D each, eps 1e-6, no time embedding; the first three end
    in Downsample2D = F.pad(x, (0,1,0,1)) + conv3x3 stride 2 padding 0); UNetMidBlock2D (resnet, single-head attention
    over H*W tokens with head_dim = 512, GroupNorm(32, eps 1e-6) on its input, residual connection, resnet);
    GroupNorm(eps 1e-6) + SiLU + conv_out 512 -> 8; quant_conv 1x1 8 -> 8 -> (mean, logvar), logvar clamped to [-30, 20].
  decoder: post_quant_conv 1x1 4 -> 4; conv_in 4 -> 512; mid block; 4 x UpDecoderBlock2D (3 ResnetBlock2D each; the first
    three end in nearest-2x Upsample2D + conv3x3); GroupNorm + SiLU + conv_out 128 -> 3.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import Config

VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, sample_size=512)


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class AttnBlock(nn.Module):
    """diffusers-0.24 Attention(_from_deprecated_attn_block=True): one head of width C, biases on q/k/v/out, softmax in
    fp32, the block input added back (residual_connection=True, rescale_output_factor=1)."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        t = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        w = torch.softmax((q @ k.transpose(1, 2)).float() / math.sqrt(C), dim=-1).to(q.dtype)
        o = self.to_out[0](w @ v)
        return o.transpose(1, 2).reshape(B, C, H, W) + x


class MidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([AttnBlock(c, groups)])
        self.resnets = nn.ModuleList([ResnetBlock(c, c, groups), ResnetBlock(c, c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Conv(nn.Module):
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=0 if stride == 2 else 1)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, groups, layers, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if down:
            self.downsamplers = nn.ModuleList([_Conv(cout, 2)])
        self.down = down

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.down:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))
        return x


class UpBlock(nn.Module):
    def __init__(self, cin, cout, groups, layers, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        if up:
            self.upsamplers = nn.ModuleList([_Conv(cout, 1)])
        self.up = up

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.up:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([DownBlock(boc[max(i - 1, 0)], c, g, cfg.layers_per_block, i < len(boc) - 1)
                                          for i, c in enumerate(boc)])
        self.mid_block = MidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = list(reversed(cfg.block_out_channels)), cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[0], 3, padding=1)
        self.mid_block = MidBlock(boc[0], g)
        self.up_blocks = nn.ModuleList([UpBlock(boc[max(i - 1, 0)], c, g, cfg.layers_per_block + 1, i < len(boc) - 1)
                                        for i, c in enumerate(boc)])
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussian:
    def __init__(self, moments):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None or
                            generator.device.type != "cpu" else "cpu", dtype=self.mean.dtype).to(self.mean.device)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class EncodeOutput:
    def __init__(self, d):
        self.latent_dist = d


class AutoencoderKL(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        self.config = Config({**VAE_CONFIG, **overrides})
        self.encoder = Encoder(self.config)
        self.decoder = Decoder(self.config)
        lc = self.config.latent_channels
        self.quant_conv = nn.Conv2d(2 * lc, 2 * lc, 1)
        self.post_quant_conv = nn.Conv2d(lc, lc, 1)

    def encode(self, x):
        return EncodeOutput(DiagonalGaussian(self.quant_conv(self.encoder(x))))

    def decode(self, z, return_dict=True, generator=None):
        out = self.decoder(self.post_quant_conv(z))
        return (out,)
