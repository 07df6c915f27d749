"""VP-SDE noise schedule and posterior buffers of GaussianDiffusion.__init__
(reference: spec_denoiser.py:26-69, diffusion_utils.py:16-18,26-45).  float64 numpy, cast to float32 at
the end — exactly what the reference does before register_buffer."""
from __future__ import annotations

import numpy as np


def get_noise_schedule_list(schedule_mode: str, timesteps: int, min_beta=0.0, max_beta=0.01, s=0.008):
    if schedule_mode == "linear":
        return np.linspace(0.000001, 0.01, timesteps)
    if schedule_mode == "cosine":
        steps = timesteps + 1
        x = np.linspace(0, steps, steps)
        ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        return np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)
    if schedule_mode == "vpsde":
        t = np.arange(1, timesteps + 1, dtype=np.float64)
        return 1.0 - np.exp(-min_beta / timesteps - 0.5 * (max_beta - min_beta) * (2 * t - 1) / timesteps ** 2)
    if schedule_mode == "logsnr":
        t = np.arange(1, timesteps + 1, dtype=np.float64) / timesteps
        b = np.arctan(np.exp(-0.5 * 20.0))
        a = np.arctan(np.exp(-0.5 * -20.0)) - b
        return -2.0 * np.log(np.tan(a * t + b))
    raise NotImplementedError(schedule_mode)


def diffusion_buffers(timesteps: int, schedule_type: str = "vpsde", betas=None) -> dict:
    """All (S+1,)-long float32 buffers the reference registers, keyed by the reference's buffer names."""
    if betas is None:
        betas = get_noise_schedule_list(schedule_type, timesteps + 1, min_beta=0.1, max_beta=40, s=0.008)
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    buf = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": np.log(np.maximum(pv, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    return {k: v.astype(np.float32) for k, v in buf.items()}
