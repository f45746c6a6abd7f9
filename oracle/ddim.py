"""Oracle: DDIMScheduler of diffusers-0.24 as the reference configures it (inference_IMAGdressing.py:119-127:
num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
set_alpha_to_one=False, steps_offset=1; default timestep_spacing="leading"). TEST INFRASTRUCTURE ONLY.

Restated from the published algorithm; known-answer values pinned in tests/test_oracle_golden.py:
timesteps(50) = 981, 961, ..., 1; timesteps(20) = 951, ..., 1; alpha-bar from the closed form.
"""
from __future__ import annotations

import numpy as np
import torch


class DDIMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.order = 1
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, x, t=None):
        return x

    def alphas(self, t: int):
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return float(a_t), float(a_p)

    def step(self, eps, t, x, eta=0.0):
        a_t, a_p = self.alphas(int(t))
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return (a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps,)

    def add_noise(self, x, noise, t):
        a = self.alphas_cumprod.to(x.device)[t].reshape(-1, 1, 1, 1).to(x.dtype)
        return a.sqrt() * x + (1 - a).sqrt() * noise
