"""DDIMScheduler with the diffusers-0.24 surface the reference touches (ctor kwargs,
`set_timesteps`, `timesteps`, `init_noise_sigma`, `order`, `scale_model_input`, `step`, `add_noise`, `config`),
configured by the reference as inference_IMAGdressing.py:119-127. Tables are computed on the host in fp32 exactly
as the published algorithm does; the update itself runs in the fused CFG+DDIM kernel."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import ops
from .modeling import FrozenConfig


class DDIMSchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if rescale_betas_zero_snr:  # train.py:403-407 (only add_noise is used with it)
            ab = torch.cumprod(1.0 - betas, 0).sqrt()
            a0, aT = ab[0].clone(), ab[-1].clone()
            ab = (ab - aT) * a0 / (a0 - aT)
            ab2 = ab ** 2
            alphas = torch.cat([ab2[0:1], ab2[1:] / ab2[:-1]])
            betas = 1 - alphas
        if clip_sample or thresholding or prediction_type != "epsilon":
            raise NotImplementedError("the reference uses clip_sample=False, epsilon prediction, no thresholding")
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                   beta_schedule=beta_schedule, clip_sample=clip_sample,
                                   set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                   prediction_type=prediction_type, timestep_spacing=timestep_spacing,
                                   rescale_betas_zero_snr=rescale_betas_zero_snr)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._dev = {}

    # ---- API
    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        elif sp == "trailing":
            ratio = T / num_inference_steps
            ts = np.round(np.arange(T, 0, -ratio)).astype(np.int64) - 1
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(sp)
        self.timesteps = torch.from_numpy(ts).to(device)

    def _alphas(self, t: int):
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_p

    def step_tables(self, device, timesteps: Optional[torch.Tensor] = None):
        """Device tables for the fused kernel: timesteps fp32 [S], coef [S,4] = sqrt(a_t), sqrt(1-a_t), sqrt(a_p),
        sqrt(1-a_p); blend [S,2] = sqrt(a_next), sqrt(1-a_next) at t_{i+1} (last row {1,0}) for the inpaint blend."""
        ts = self.timesteps if timesteps is None else timesteps
        key = (str(device), tuple(int(t) for t in ts))
        hit = self._dev.get(key)
        if hit is None:
            tl = [int(t) for t in ts]
            coef, blend = [], []
            for i, t in enumerate(tl):
                a_t, a_p = self._alphas(t)
                coef.append([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)])
                if i + 1 < len(tl):
                    a_n = float(self.alphas_cumprod[tl[i + 1]])
                    blend.append([math.sqrt(a_n), math.sqrt(1 - a_n)])
                else:
                    blend.append([1.0, 0.0])
            hit = (torch.tensor(tl, dtype=torch.float32, device=device), torch.tensor(coef, dtype=torch.float32, device=device),
                   torch.tensor(blend, dtype=torch.float32, device=device))
            self._dev[key] = hit
        return hit

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """x_{t-1} from eps (eta = 0, deterministic). CUDA tensors go through the fused kernel."""
        if eta != 0.0:
            raise NotImplementedError("every reference script samples with eta = 0")
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        a_t, a_p = self._alphas(int(timestep))
        x = sample.float().contiguous()
        out = x.clone()
        coef = torch.tensor([[math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)]],
                            dtype=torch.float32, device=x.device)
        step = torch.zeros(2, dtype=torch.int32, device=x.device)
        ops.cfg_ddim_step(model_output.float().contiguous(), None, 1.0, out, coef, step)
        out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return DDIMSchedulerOutput(out)

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps].to(original_samples.dtype)
        while a.dim() < original_samples.dim():
            a = a.unsqueeze(-1)
        return a.sqrt() * original_samples + (1 - a).sqrt() * noise

    @classmethod
    def from_config(cls, config, **kw):
        return cls(**{**{k: v for k, v in dict(config).items() if k in cls.__init__.__code__.co_varnames}, **kw})
