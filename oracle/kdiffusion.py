"""ORACLE (test infrastructure) — restatement of the k-diffusion pieces the webui calls on the hot path.

Upstream (un-vendored, crowsonkb/k-diffusion@ab527a9a, pinned modules/launch_utils.py:357):
  k_diffusion/external.py  DiscreteSchedule, DiscreteEpsDDPMDenoiser, CompVisDenoiser
  k_diffusion/sampling.py  get_sigmas_karras, get_ancestral_step, sample_euler_ancestral, sample_dpmpp_2m
In-tree call sites / anchors: modules/sd_samplers_kdiffusion.py:61-62 (CompVisDenoiser(sd_model, quantize=False)),
:79-132 (get_sigmas), :190-234 (sample), :134-188 (sample_img2img); modules/sd_schedulers.py:10-15 (to_d override);
modules/sd_samplers_common.py:205-226 (randn_like -> p.rng.next()); partial twin modules/sd_samplers_lcm.py:35-63;
beta schedule modules/models/diffusion/ddpm_edit.py:133-141 + configs/v1-inference.yaml:5-9.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def make_alphas_cumprod(linear_start=0.00085, linear_end=0.012, timesteps=1000) -> torch.Tensor:
    """ldm make_beta_schedule("linear") — 'scaled linear' in sqrt space (ddpm_edit.py:133-141; float64 then fp32)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    alphas = 1.0 - betas
    return torch.tensor(np.cumprod(alphas, axis=0), dtype=torch.float32)


class DiscreteSchedule:
    """k_diffusion.external.DiscreteSchedule (quantize=False as the webui constructs it)."""

    def __init__(self, alphas_cumprod: torch.Tensor):
        self.sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
        self.log_sigmas = self.sigmas.log()

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def get_sigmas(self, n: int) -> torch.Tensor:
        """linspace over the discrete timesteps, append zero ('Automatic' scheduler for Euler a)."""
        t_max = len(self.sigmas) - 1
        t = torch.linspace(t_max, 0, n, device=self.sigmas.device)
        return torch.cat([self.t_to_sigma(t), t.new_zeros([1])])

    def sigma_to_t(self, sigma: torch.Tensor) -> torch.Tensor:
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas.to(sigma.device)[:, None]
        low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        ls = self.log_sigmas.to(sigma.device)
        low, high = ls[low_idx], ls[high_idx]
        w = ((low - log_sigma) / (low - high)).clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def t_to_sigma(self, t: torch.Tensor) -> torch.Tensor:
        t = t.float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        ls = self.log_sigmas.to(t.device)
        log_sigma = (1 - w) * ls[low_idx] + w * ls[high_idx]
        return log_sigma.exp()


class CompVisDenoiser(DiscreteSchedule):
    """k_diffusion.external.CompVisDenoiser / DiscreteEpsDDPMDenoiser: eps-prediction wrapper.

    `apply_model(x, t, cond)` is the (patched) LatentDiffusion.apply_model — modules/sd_hijack_unet.py:40-54 casts
    x, t and cond to dtype_unet first (so the fractional timestep is rounded to fp16 on the fp16 path)."""

    sigma_data = 1.0

    def __init__(self, apply_model, alphas_cumprod):
        super().__init__(alphas_cumprod)
        self.apply_model = apply_model

    def get_scalings(self, sigma):
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_out, c_in

    def __call__(self, x, sigma, **kwargs):
        c_out, c_in = [s.view(-1, 1, 1, 1) for s in self.get_scalings(sigma)]
        eps = self.apply_model(x * c_in, self.sigma_to_t(sigma), **kwargs)
        return x + eps * c_out


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu") -> torch.Tensor:
    ramp = torch.linspace(0, 1, n, device=device)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def to_d(x, sigma, denoised):
    """modules/sd_schedulers.py:10-15 (the webui's override: plain division, no append_dims)."""
    return (x - denoised) / sigma


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


@torch.no_grad()
def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, eta=1.0, s_noise=1.0, noise_sampler=None):
    """noise_sampler() is `p.rng.next()` in the webui (TorchHijack.randn_like, sd_samplers_common.py:225-226)."""
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        d = to_d(x, sigmas[i], denoised)
        dt = sigma_down - sigmas[i]
        x = x + d * dt
        if sigmas[i + 1] > 0:
            x = x + noise_sampler() * s_noise * sigma_up
    return x


@torch.no_grad()
def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])

    def sigma_fn(t):
        return t.neg().exp()

    def t_fn(sigma):
        return sigma.log().neg()

    old_denoised = None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_denoised is None or sigmas[i + 1] == 0:
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_d
        old_denoised = denoised
    return x


def setup_img2img_steps(steps_requested: int, denoising_strength: float, steps_given: bool = True):
    """modules/sd_samplers_common.py:22-31. `steps_given` = the caller passed `steps` (the hires second pass does,
    :25-27); plain img2img does not and, with img2img_fix_steps off (the default), takes the else branch (:28-30)."""
    if steps_given:
        steps = int(steps_requested / min(denoising_strength, 0.999)) if denoising_strength > 0 else 0
        t_enc = steps_requested - 1
    else:
        steps = steps_requested
        t_enc = int(min(denoising_strength, 0.999) * steps)
    return steps, t_enc


def webui_sigmas(schedule: DiscreteSchedule, sampler: str, steps: int) -> torch.Tensor:
    """modules/sd_samplers_kdiffusion.py:79-132 with default options: 'Euler a' -> model_wrap.get_sigmas(steps);
    'DPM++ 2M' -> its default scheduler 'karras' with the model's sigma_min / sigma_max, rho 7. CPU tensor."""
    if sampler == "Euler a":
        return schedule.get_sigmas(steps).cpu()
    if sampler in ("DPM++ 2M", "DPM++ 2M Karras"):
        return get_sigmas_karras(steps, schedule.sigmas[0].item(), schedule.sigmas[-1].item(), 7.0, "cpu")
    raise ValueError(sampler)
