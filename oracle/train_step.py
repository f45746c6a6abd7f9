"""Oracle restatement of the reference's training step (BASELINE.json configs[4]).
TEST INFRASTRUCTURE ONLY — plain PyTorch, autograd supplies the backward pass; the product has no training kernels yet.

Follows /root/reference/train.py:
  * compute_snr                     train.py:214-241
  * SDModel.forward                 train.py:255-281   (garment pass WITH grad on the full batch, every cache row kept —
                                                        unlike inference, which keeps row [1] of a batch-2 call)
  * processor / trainable set-up    train.py:338-376   (to_k_ref / to_v_ref start as copies of to_k / to_v; trainable =
                                                        image projection + garment UNet + adapter modules; denoising UNet frozen)
  * noise / timesteps / loss        train.py:527-605   (epsilon target; plain MSE or min-SNR-gamma weighting)
Parity unpinned (the reference has no training tests or fixtures); known-answer checks are in tests/test_train_oracle.py.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.nn.functional as F

from . import processors as op


def compute_snr(alphas_cumprod: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """SNR(t) = (sqrt(abar_t) / sqrt(1 - abar_t))^2   (train.py:214-241)."""
    alpha = (alphas_cumprod ** 0.5).to(timesteps.device)[timesteps].float()
    sigma = ((1.0 - alphas_cumprod) ** 0.5).to(timesteps.device)[timesteps].float()
    return (alpha / sigma) ** 2


def hidden_size_of(name: str, block_out_channels) -> int:
    """train.py:341-348 (same rule as inference_IMAGdressing.py:72-79)."""
    if name.startswith("mid_block"):
        return block_out_channels[-1]
    if name.startswith("up_blocks"):
        return list(reversed(block_out_channels))[int(name[len("up_blocks.")])]
    return block_out_channels[int(name[len("down_blocks.")])]


def install_training_processors(unet, ref_unet) -> torch.nn.ModuleList:
    """train.py:338-366: RefS processors on attn1 (to_k_ref / to_v_ref initialised from the layer's own to_k / to_v),
    C processors on attn2, cache processors on the garment UNet. Returns `adapter_modules`."""
    st = unet.state_dict()
    procs = {}
    for name in unet.attn_processors.keys():
        hidden = hidden_size_of(name, unet.config.block_out_channels)
        if name.endswith("attn1.processor"):
            p = op.RefSAttnProcessor(name, hidden)
            layer = name.split(".processor")[0]
            p.load_state_dict({"to_k_ref.weight": st[layer + ".to_k.weight"], "to_v_ref.weight": st[layer + ".to_v.weight"]})
            procs[name] = p
        else:
            procs[name] = op.CAttnProcessor(name, hidden, unet.config.cross_attention_dim)
    unet.set_attn_processor(procs)
    ref_unet.set_attn_processor({n: op.CacheAttnProcessor() for n in ref_unet.attn_processors.keys()})
    return torch.nn.ModuleList(unet.attn_processors.values())


def set_trainable(unet, ref_unet, proj, adapter_modules) -> Iterable[torch.nn.Parameter]:
    """train.py:368-379. Note the ORDER: unet.requires_grad_(False) also freezes the adapter modules registered inside it;
    adapter_modules.requires_grad_(True) re-enables exactly those."""
    unet.requires_grad_(False)
    proj.requires_grad_(True)
    ref_unet.requires_grad_(True)
    adapter_modules.requires_grad_(True)
    return [*proj.parameters(), *ref_unet.parameters(), *adapter_modules.parameters()]


def sd_forward(unet, ref_unet, proj, encoder_hidden_states, latents, ref_latents, clip_image_embeddings, timesteps):
    """SDModel.forward (train.py:255-281): noise prediction of the denoising UNet conditioned on the garment taps."""
    cloth = proj(clip_image_embeddings)                                  # :257
    ref_unet(ref_latents, torch.zeros_like(timesteps), cloth)            # :259-264 (output discarded)
    sa: Dict[str, torch.Tensor] = {n: p.cache["hidden_states"] for n, p in ref_unet.attn_processors.items()}  # :266-268
    return unet(latents, timesteps, encoder_hidden_states, cross_attention_kwargs={"sa_hidden_states": sa})[0]  # :272-279


def training_loss(model_pred, target, alphas_cumprod: Optional[torch.Tensor] = None, timesteps=None, snr_gamma: float = 0.0):
    """train.py:573-596 for prediction_type == "epsilon"."""
    if snr_gamma == 0:
        return F.mse_loss(model_pred.float(), target.float(), reduction="mean")
    snr = compute_snr(alphas_cumprod, timesteps)
    w = torch.stack([snr, snr_gamma * torch.ones_like(timesteps)], dim=1).min(dim=1)[0] / snr
    loss = F.mse_loss(model_pred.float(), target.float(), reduction="none")
    return (loss.mean(dim=list(range(1, loss.dim()))) * w).mean()


def train_step(unet, ref_unet, proj, scheduler, latents, ref_latents, clip_image_embeddings, encoder_hidden_states,
               noise, timesteps, snr_gamma: float = 0.0):
    """One micro-batch of train.py:527-600 after the (frozen, excluded) VAE / CLIP encoders: add noise, predict, loss,
    backward. Returns the loss; gradients are left on the trainable parameters."""
    noisy = scheduler.add_noise(latents, noise, timesteps)               # :545
    pred = sd_forward(unet, ref_unet, proj, encoder_hidden_states, noisy, ref_latents, clip_image_embeddings, timesteps)
    loss = training_loss(pred, noise, scheduler.alphas_cumprod, timesteps, snr_gamma)  # target = noise (:562)
    loss.backward()                                                      # :603
    return loss.detach()
