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


# ----------------------------------------------------------------------------------------------------------------------
# The other k-diffusion samplers the webui lists (modules/sd_samplers_kdiffusion.py:11-27), restated from
# crowsonkb/k-diffusion@ab527a9a k_diffusion/sampling.py (un-vendored). noise_sampler() = p.rng.next() as above.
# ----------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_euler(model, x, sigmas, extra_args=None, callback=None, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_sampler=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + noise_sampler() * s_noise * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * s_in, **extra_args)
        d = to_d(x, sigma_hat, denoised)
        x = x + d * (sigmas[i + 1] - sigma_hat)
    return x


@torch.no_grad()
def sample_heun(model, x, sigmas, extra_args=None, callback=None, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_sampler=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + noise_sampler() * s_noise * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * s_in, **extra_args)
        d = to_d(x, sigma_hat, denoised)
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            denoised_2 = model(x_2, sigmas[i + 1] * s_in, **extra_args)
            d_2 = to_d(x_2, sigmas[i + 1], denoised_2)
            x = x + (d + d_2) / 2 * dt
    return x


@torch.no_grad()
def sample_dpm_2(model, x, sigmas, extra_args=None, callback=None, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, noise_sampler=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            x = x + noise_sampler() * s_noise * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(x, sigma_hat * s_in, **extra_args)
        d = to_d(x, sigma_hat, denoised)
        if sigmas[i + 1] == 0:
            x = x + d * (sigmas[i + 1] - sigma_hat)
        else:
            sigma_mid = sigma_hat.log().lerp(sigmas[i + 1].log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigma_hat)
            denoised_2 = model(x_2, sigma_mid * s_in, **extra_args)
            d_2 = to_d(x_2, sigma_mid, denoised_2)
            x = x + d_2 * (sigmas[i + 1] - sigma_hat)
    return x


@torch.no_grad()
def sample_dpm_2_ancestral(model, x, sigmas, extra_args=None, callback=None, eta=1.0, s_noise=1.0, noise_sampler=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        d = to_d(x, sigmas[i], denoised)
        if sigma_down == 0:
            x = x + d * (sigma_down - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigma_down.log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigmas[i])
            denoised_2 = model(x_2, sigma_mid * s_in, **extra_args)
            d_2 = to_d(x_2, sigma_mid, denoised_2)
            x = x + d_2 * (sigma_down - sigmas[i])
            x = x + noise_sampler() * s_noise * sigma_up
    return x


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, x, sigmas, extra_args=None, callback=None, eta=1.0, s_noise=1.0, noise_sampler=None):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()  # noqa: E731
    t_fn = lambda sigma: sigma.log().neg()  # noqa: E731
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if sigma_down == 0:
            d = to_d(x, sigmas[i], denoised)
            x = x + d * (sigma_down - sigmas[i])
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            x_2 = (sigma_fn(s) / sigma_fn(t)) * x - (-h * r).expm1() * denoised
            denoised_2 = model(x_2, sigma_fn(s) * s_in, **extra_args)
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * denoised_2
        if sigmas[i + 1] > 0:
            x = x + noise_sampler() * s_noise * sigma_up
    return x


def linear_multistep_coeff(order, t, i, j):
    from scipy import integrate

    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")

    def fn(tau):
        prod = 1.0
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod

    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


@torch.no_grad()
def sample_lms(model, x, sigmas, extra_args=None, callback=None, order=4):
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sigmas_cpu = sigmas.detach().cpu().numpy()
    ds = []
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        d = to_d(x, sigmas[i], denoised)
        ds.append(d)
        if len(ds) > order:
            ds.pop(0)
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sigmas_cpu, i, j) for j in range(cur_order)]
        x = x + sum(coeff * d for coeff, d in zip(coeffs, reversed(ds)))
    return x


def get_sigmas_exponential(n, sigma_min, sigma_max, device="cpu"):
    sigmas = torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp()
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0, device="cpu"):
    ramp = torch.linspace(1, 0, n, device=device) ** rho
    sigmas = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
    return torch.cat([sigmas, sigmas.new_zeros([1])])


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


SAMPLER_TABLE = {  # label -> (function name, default scheduler, discard_next_to_last_sigma); sd_samplers_kdiffusion.py:11-27
    "Euler a": ("sample_euler_ancestral", None, False), "DPM++ 2M": ("sample_dpmpp_2m", "karras", False),
    "DPM++ 2M Karras": ("sample_dpmpp_2m", "karras", False), "Euler": ("sample_euler", None, False),
    "Heun": ("sample_heun", None, False), "LMS": ("sample_lms", None, False), "DPM2": ("sample_dpm_2", "karras", True),
    "DPM2 a": ("sample_dpm_2_ancestral", "karras", True), "DPM++ 2S a": ("sample_dpmpp_2s_ancestral", "karras", False),
}


def webui_sigmas(schedule: DiscreteSchedule, sampler: str, steps: int) -> torch.Tensor:
    """modules/sd_samplers_kdiffusion.py:79-132 with default options: samplers without a default scheduler ->
    model_wrap.get_sigmas(steps); 'karras' -> the model's sigma_min / sigma_max, rho 7; discard_next_to_last_sigma samplers
    ask for one more step and drop the penultimate sigma. CPU tensor."""
    _, sched, discard = SAMPLER_TABLE[sampler]
    n = steps + (1 if discard else 0)
    if sched is None:
        sig = schedule.get_sigmas(n).cpu()
    else:
        sig = get_sigmas_karras(n, schedule.sigmas[0].item(), schedule.sigmas[-1].item(), 7.0, "cpu")
    if discard:
        sig = torch.cat([sig[:-2], sig[-1:]])
    return sig
