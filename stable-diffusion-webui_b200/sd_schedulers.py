"""Noise schedules: host mirror of modules/sd_schedulers.py (same names, arguments and registry), plus the three schedules
that file takes from k-diffusion (karras, exponential, polyexponential; k_diffusion/sampling.py, un-vendored).

Every function returns n + 1 sigmas on `device` (the last one 0). `inner_model` is the CompVisDenoiser wrapper
(sigmas table, sigma_to_t, t_to_sigma, get_sigmas). The in-tree functions are pinned to the reference's own code on the
same inputs (tests/golden/make_golden_sched.py -> tests/test_schedulers_cpu.py).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, List, Optional

import numpy as np
import torch


@dataclasses.dataclass
class Scheduler:
    name: str
    label: str
    function: Any
    default_rho: float = -1
    need_inner_model: bool = False
    aliases: Optional[List[str]] = None


def uniform(n, sigma_min, sigma_max, inner_model, device):
    return inner_model.get_sigmas(n).to(device)


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    ramp = torch.linspace(0, 1, n, device=device)
    min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def get_sigmas_exponential(n, sigma_min, sigma_max, device="cpu"):
    sigmas = torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp()
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0, device="cpu"):
    ramp = torch.linspace(1, 0, n, device=device) ** rho
    sigmas = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def sgm_uniform(n, sigma_min, sigma_max, inner_model, device):
    start = inner_model.sigma_to_t(torch.tensor(sigma_max))
    end = inner_model.sigma_to_t(torch.tensor(sigma_min))
    sigs = [inner_model.t_to_sigma(ts) for ts in torch.linspace(start, end, n + 1)[:-1]]
    return torch.FloatTensor(sigs + [0.0]).to(device)


AYS_SD15 = [14.615, 6.475, 3.861, 2.697, 1.886, 1.396, 0.963, 0.652, 0.399, 0.152, 0.029]
AYS_SDXL = [14.615, 6.315, 3.771, 2.181, 1.342, 0.862, 0.555, 0.380, 0.234, 0.113, 0.029]


def get_align_your_steps_sigmas(n, sigma_min, sigma_max, device, is_sdxl: bool = False):
    """https://research.nvidia.com/labs/toronto-ai/AlignYourSteps/howto.html — the 10-step table, log-linearly
    interpolated to n points when n != 11 (the reference reads `shared.sd_model.is_sdxl`; here it is an argument)."""
    table = list(AYS_SDXL if is_sdxl else AYS_SD15)
    if n != len(table):
        xs = np.linspace(0, 1, len(table))
        ys = np.log(table[::-1])
        sig = np.exp(np.interp(np.linspace(0, 1, n), xs, ys))[::-1].copy()
        sig = np.append(sig, [0.0])
    else:
        sig = table + [0.0]
    return torch.FloatTensor(sig).to(device)


def kl_optimal(n, sigma_min, sigma_max, device):
    alpha_min = torch.arctan(torch.tensor(sigma_min, device=device))
    alpha_max = torch.arctan(torch.tensor(sigma_max, device=device))
    step = torch.arange(n + 1, device=device)
    return torch.tan(step / n * alpha_min + (1.0 - step / n) * alpha_max)


def simple_scheduler(n, sigma_min, sigma_max, inner_model, device):
    ss = len(inner_model.sigmas) / n
    sigs = [float(inner_model.sigmas[-(1 + int(x * ss))]) for x in range(n)]
    return torch.FloatTensor(sigs + [0.0]).to(device)


def normal_scheduler(n, sigma_min, sigma_max, inner_model, device, sgm=False, floor=False):
    start = inner_model.sigma_to_t(torch.tensor(sigma_max))
    end = inner_model.sigma_to_t(torch.tensor(sigma_min))
    timesteps = torch.linspace(start, end, n + 1)[:-1] if sgm else torch.linspace(start, end, n)
    sigs = [inner_model.t_to_sigma(ts) for ts in timesteps]
    return torch.FloatTensor(sigs + [0.0]).to(device)


def ddim_scheduler(n, sigma_min, sigma_max, inner_model, device):
    ss = max(len(inner_model.sigmas) // n, 1)
    sigs = [float(inner_model.sigmas[x]) for x in range(1, len(inner_model.sigmas), ss)]
    return torch.FloatTensor(sigs[::-1] + [0.0]).to(device)


def beta_scheduler(n, sigma_min, sigma_max, inner_model, device, alpha: float = 0.6, beta: float = 0.6):
    """"Beta Sampling is All You Need" (arXiv:2407.12173); alpha / beta are opts.beta_dist_alpha / beta_dist_beta (0.6, 0.6)."""
    from scipy import stats

    ts = [stats.beta.ppf(x, alpha, beta) for x in 1 - np.linspace(0, 1, n)]
    return torch.FloatTensor([sigma_min + (x * (sigma_max - sigma_min)) for x in ts] + [0.0]).to(device)


schedulers = [
    Scheduler("automatic", "Automatic", None),
    Scheduler("uniform", "Uniform", uniform, need_inner_model=True),
    Scheduler("karras", "Karras", get_sigmas_karras, default_rho=7.0),
    Scheduler("exponential", "Exponential", get_sigmas_exponential),
    Scheduler("polyexponential", "Polyexponential", get_sigmas_polyexponential, default_rho=1.0),
    Scheduler("sgm_uniform", "SGM Uniform", sgm_uniform, need_inner_model=True, aliases=["SGMUniform"]),
    Scheduler("kl_optimal", "KL Optimal", kl_optimal),
    Scheduler("align_your_steps", "Align Your Steps", get_align_your_steps_sigmas),
    Scheduler("simple", "Simple", simple_scheduler, need_inner_model=True),
    Scheduler("normal", "Normal", normal_scheduler, need_inner_model=True),
    Scheduler("ddim", "DDIM", ddim_scheduler, need_inner_model=True),
    Scheduler("beta", "Beta", beta_scheduler, need_inner_model=True),
]
schedulers_map = {**{x.name: x for x in schedulers}, **{x.label: x for x in schedulers}}
