"""DDIM scheduler (eta = 0) with the SD-1.5 noise schedule -- host-side scalars only.

BASELINE.json measures "50-step DDIM generation"; the reference itself samples with
DPMSolverMultistepScheduler (models/infer.py:39-40) built from the SD-1.5 DDPM config
(1000 steps, scaled_linear betas 0.00085 -> 0.012, epsilon prediction, steps_offset=1).
Both arms of every parity / benchmark comparison use this same scheduler, so it does not enter parity.

All coefficients are computed once in float64 on the host; the per-step update is
    x_{t-1} = c_x[i] * x_t + c_eps[i] * eps_theta
so the denoise loop performs no device->host synchronisation.
"""
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class DDIMSchedule:
    timesteps: List[int]      # descending, length = num_inference_steps
    c_x: List[float]
    c_eps: List[float]
    init_noise_sigma: float = 1.0


def make_ddim_schedule(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                       beta_end: float = 0.012, steps_offset: int = 1, set_alpha_to_one: bool = False) -> DDIMSchedule:
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    alphas_cumprod = np.cumprod(1.0 - betas)
    final_alpha = 1.0 if set_alpha_to_one else alphas_cumprod[0]
    ratio = num_train_timesteps // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].astype(np.int64) + steps_offset   # "leading"
    c_x, c_eps = [], []
    for t in ts:
        prev = t - ratio
        a_t = alphas_cumprod[t]
        a_prev = alphas_cumprod[prev] if prev >= 0 else final_alpha
        # x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t);  x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps
        c_x.append(float(np.sqrt(a_prev / a_t)))
        c_eps.append(float(np.sqrt(1.0 - a_prev) - np.sqrt(a_prev * (1.0 - a_t) / a_t)))
    return DDIMSchedule([int(t) for t in ts], c_x, c_eps)
