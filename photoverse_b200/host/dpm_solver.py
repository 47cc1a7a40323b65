"""DPM-Solver++(2M) multistep scheduler -- the sampler the reference actually uses
(``DPMSolverMultistepScheduler.from_config(...)``, models/infer.py:39-40; diffusers is not installable here, so this is
a restatement of the published algorithm: Lu et al., "DPM-Solver++", 2022, Alg. 2, data-prediction form, with the
SD-1.5 noise schedule of the reference's ``DDPMScheduler`` config: 1000 steps, scaled_linear betas 0.00085 -> 0.012).

Host-side scalars only.  Every step is an affine combination of tensors the loop already holds,

    x0_i   = (x_i - sigma_i * eps_i) / alpha_i                      (epsilon-prediction UNet -> data prediction)
    D_i    = x0_i                                   first step / last step
           = (1 + 1/(2 r_i)) x0_i - 1/(2 r_i) x0_{i-1}              r_i = h_{i-1} / h_i     (2M)
    x_{i+1} = (sigma_{i+1} / sigma_i) x_i - alpha_{i+1} (exp(-h_i) - 1) D_i ,     h_i = lambda_{i+1} - lambda_i

The last step follows diffusers 0.27.2's defaults (``final_sigmas_type="zero"``, ``lower_order_final=True``): the sigma
grid is closed with sigma = 0 (alpha = 1), which makes the final update first-order for every N and lands exactly on the
data prediction, x_N = x0_{N-1} [3P-recall of scheduling_dpmsolver_multistep.py ``set_timesteps`` / ``step``].
The denoise loop stays free of device->host synchronisation.  BASELINE.json measures DDIM (host/ddim.py); this is
the "next" row f1 of SURVEY.md section 8.
"""
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class DPMSchedule:
    timesteps: List[int]          # descending, length N
    # x_next = cx[i] * x + c0[i] * x0_i + c0p[i] * x0_{i-1};   x0_i = kx[i] * x + ke[i] * eps_i
    cx: List[float]
    c0: List[float]
    c0p: List[float]
    kx: List[float]
    ke: List[float]
    init_noise_sigma: float = 1.0


def make_dpmpp_2m_schedule(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                           beta_end: float = 0.012, final_sigmas_type: str = "zero") -> DPMSchedule:
    """``final_sigmas_type``: "zero" (diffusers default: terminal sigma 0) or "sigma_min" (terminal node t = 0, the last
    step first-order only for N < 15 -- diffusers' ``lower_order_final`` rule for that setting)."""
    if final_sigmas_type not in ("zero", "sigma_min"):
        raise ValueError("final_sigmas_type must be 'zero' or 'sigma_min'")
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    # "linspace" timestep spacing over [0, T-1], descending; the terminal node closes the grid
    ts = np.linspace(0, num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1].astype(np.int64)
    alpha = lambda t: np.sqrt(ac[t])
    sigma = lambda t: np.sqrt(1.0 - ac[t])
    lam = lambda t: np.log(alpha(t) / sigma(t))
    nodes = list(ts) + [0]
    zero_final = final_sigmas_type == "zero"
    cx, c0, c0p, kx, ke = [], [], [], [], []
    for i in range(num_inference_steps):
        s, t = nodes[i], nodes[i + 1]
        first = i == 0
        last = i == num_inference_steps - 1
        if last and zero_final:
            # sigma_next = 0, alpha_next = 1, h = +inf: x_next = x0_i exactly
            cx.append(0.0); c0.append(1.0); c0p.append(0.0)
            kx.append(float(1.0 / alpha(s)))
            ke.append(float(-sigma(s) / alpha(s)))
            continue
        h = lam(t) - lam(s)
        a = sigma(t) / sigma(s)
        b = -alpha(t) * (np.exp(-h) - 1.0)
        last_lower = last and num_inference_steps < 15
        if first or last_lower:
            c0.append(float(b)); c0p.append(0.0)
        else:
            h_prev = lam(s) - lam(nodes[i - 1])
            r = h_prev / h
            c0.append(float(b * (1.0 + 1.0 / (2.0 * r)))); c0p.append(float(-b / (2.0 * r)))
        cx.append(float(a))
        kx.append(float(1.0 / alpha(s)))
        ke.append(float(-sigma(s) / alpha(s)))
    return DPMSchedule([int(t) for t in ts], cx, c0, c0p, kx, ke)
