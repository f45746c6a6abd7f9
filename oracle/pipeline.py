"""Oracle restatement of the reference denoising loop (dressing_sd/pipelines/IMAGDressing_v1_pipeline.py:463-541;
ControlNet variant IMAGDressing_v1_pipeline_ipa_controlnet.py:595-736; inpainting blend
IMAGDressing_v1_pipeline_controlnet_inpainting.py:487-500) at batch 1, fp32, two separate UNet calls per step
exactly as the reference does. TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch

from .ddim import DDIMOracle


@torch.no_grad()
def sample_one(unet, ref_unet, latents, prompt_embeds, negative_embeds, garment_tokens, ref_latents, guidance, steps,
               controlnet=None, control_cond=None, control_scale=1.0, control_text=None, mask=None,
               image_latents=None, noise=None, control_guidance_start=0.0, control_guidance_end=1.0):
    """One image. latents [1,4,h,w]; embeds [1,T,768]; garment_tokens [1,16,768]; returns final latents.
    control_guidance_start / _end: the `controlnet_keep` window (ipa_controlnet.py:584-590, applied :643-649)."""
    sch = DDIMOracle()
    sch.set_timesteps(steps, device=latents.device)
    ts = sch.timesteps
    sa = None
    for i, t in enumerate(ts):
        if i == 0:  # :465-479 — garment pass at t = 0, keep the attn1 processor inputs
            ref_unet(ref_latents, torch.zeros_like(t), garment_tokens)
            sa = {n: p.cache["hidden_states"] for n, p in ref_unet.attn_processors.items()}
        down_c = mid_c = down_u = mid_u = None
        if controlnet is not None:  # ipa_controlnet.py:651-666 — batch-2 call, [uncond, cond] text
            ct_c, ct_u = control_text if control_text is not None else (prompt_embeds, negative_embeds)
            keep = 1.0 - float(i / len(ts) < control_guidance_start or (i + 1) / len(ts) > control_guidance_end)
            down, mid = controlnet(torch.cat([latents] * 2), t, torch.cat([ct_u, ct_c]), control_cond,
                                   conditioning_scale=control_scale * keep)
            down_c, mid_c = [d[1:2] for d in down], mid[1:2]
            down_u, mid_u = [d[0:1] for d in down], mid[0:1]
        eps_c = unet(latents, t, prompt_embeds, cross_attention_kwargs={"sa_hidden_states": sa},
                     down_block_additional_residuals=down_c, mid_block_additional_residual=mid_c)[0]  # :499-509
        eps_u = unet(latents, t, negative_embeds, down_block_additional_residuals=down_u,
                     mid_block_additional_residual=mid_u)[0]  # :511-518 (no garment stream)
        eps = eps_u + guidance * (eps_c - eps_u)  # :521-527
        latents = sch.step(eps, t, latents)[0]  # :530-532
        if mask is not None:  # inpainting.py:487-500
            proper = image_latents
            if i < len(ts) - 1:
                proper = sch.add_noise(image_latents, noise, ts[i + 1:i + 2])
            latents = (1 - mask) * proper + mask * latents
    return latents
