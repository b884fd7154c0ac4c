"""Noise schedules and the 13 fp32 `[T]` buffers of `GaussianDiffusion` (ddpm.py:460-494, 547-615).

Derived in fp64 with torch and cast to fp32 exactly once, like the reference, so that the values
are bit-identical to the reference's registered buffers (tests/test_schedule.py).
"""
import math
from collections import OrderedDict

import torch

BUFFER_NAMES = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
    "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2", "loss_weight",
)


def betas_for(name: str, timesteps: int, **kw) -> torch.Tensor:
    f64 = torch.float64
    if name == "linear":
        k = 1000 / timesteps
        return torch.linspace(k * 0.0001, k * 0.02, timesteps, dtype=f64)
    grid = torch.linspace(0, timesteps, timesteps + 1, dtype=f64) / timesteps
    if name == "cosine":
        s = kw.get("s", 0.008)
        abar = torch.cos((grid + s) / (1 + s) * math.pi * 0.5) ** 2
    elif name == "sigmoid":
        start, end, tau = kw.get("start", -3), kw.get("end", 3), kw.get("tau", 1)
        lo, hi = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
        abar = (hi - ((grid * (end - start) + start) / tau).sigmoid()) / (hi - lo)
    else:
        raise ValueError(f"unknown beta schedule {name}")
    abar = abar / abar[0]
    return torch.clip(1 - abar[1:] / abar[:-1], 0, 0.999)


def make_buffers(name: str, timesteps: int, objective: str, min_snr_loss_weight=False, min_snr_gamma=5, **kw):
    beta = betas_for(name, timesteps, **kw)
    alpha = 1.0 - beta
    abar = torch.cumprod(alpha, dim=0)
    abar_prev = torch.cat([torch.ones(1, dtype=abar.dtype), abar[:-1]])
    post_var = beta * (1.0 - abar_prev) / (1.0 - abar)
    snr = abar / (1 - abar)
    csnr = snr.clone()
    if min_snr_loss_weight:
        csnr.clamp_(max=min_snr_gamma)
    lw = {"pred_noise": csnr / snr, "pred_x0": csnr, "pred_v": csnr / (snr + 1)}[objective]
    vals = (
        beta, abar, abar_prev, torch.sqrt(abar), torch.sqrt(1.0 - abar), torch.log(1.0 - abar),
        torch.sqrt(1.0 / abar), torch.sqrt(1.0 / abar - 1), post_var, torch.log(post_var.clamp(min=1e-20)),
        beta * torch.sqrt(abar_prev) / (1.0 - abar), (1.0 - abar_prev) * torch.sqrt(alpha) / (1.0 - abar), lw,
    )
    return OrderedDict((n, v.to(torch.float32)) for n, v in zip(BUFFER_NAMES, vals))
