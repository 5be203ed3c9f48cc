"""ORACLE (test infrastructure) — the reference's txt2img / hires-fix sampling path, restated end to end.

Follows modules/processing.py:863-1091 (process_images_inner) for what touches the GPU on the hot path:
  ImageRNG per batch (:949) -> Txt2Img.sample (:1307-1362): x = rng.next(); sampler.sample(...)
  -> [hires: sample_hr_pass :1364-1463, latent upscale F.interpolate(mode=bilinear, antialias False) :1392,
      new ImageRNG :1429, sampler.sample_img2img :1454]
  -> decode_latent_batch (:625-672): ONE image at a time, z / scale_factor, AutoencoderKL.decode, no autocast
  -> clamp((x + 1) / 2, 0, 1) (:1004-1005).
Sampler glue: modules/sd_samplers_kdiffusion.py:190-234 (x * sigmas[0]) and :134-188 (img2img).

`device="cpu", dtype=float32` is BASELINE config 1 (`--use-cpu all --no-half`); `device="cuda", dtype=float16,
autocast=True` is the reference's default-SDP GPU path.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import kdiffusion as K
from .cfg_denoiser import CFGDenoiser, make_apply_model
from .rng import ImageRNG


@dataclass
class SamplingParams:
    sampler: str = "Euler a"         # or "DPM++ 2M" (its default scheduler is karras)
    steps: int = 20
    cfg_scale: float = 7.0
    width: int = 512
    height: int = 512
    seeds: tuple = (1000,)
    randn_source: str = "NV"         # "NV" = CPU Philox (reproducible anywhere); "GPU" = torch CUDA generators
    scale_factor: float = 0.18215    # SD1.5; SDXL 0.13025
    enable_hr: bool = False
    hr_scale: float = 2.0
    hr_second_pass_steps: int = 0
    denoising_strength: float = 0.75
    eta: float = 1.0
    s_noise: float = 1.0


class OraclePipeline:
    def __init__(self, unet, vae, device, dtype_unet=torch.float32, dtype_vae=None, autocast=False):
        self.unet, self.vae = unet, vae
        self.device = torch.device(device)
        self.dtype_unet = dtype_unet
        self.dtype_vae = dtype_vae or dtype_unet
        self.autocast = autocast
        self.alphas_cumprod = K.make_alphas_cumprod().to(self.device)
        self.model_wrap = K.CompVisDenoiser(make_apply_model(unet, dtype_unet, autocast), self.alphas_cumprod)

    # -- sampler -------------------------------------------------------------------------------------------------
    def _run_sampler(self, p: SamplingParams, x, sigmas, rng, cond, uncond, y_cond, y_uncond):
        cfg = CFGDenoiser(self.model_wrap)
        extra = {"cond": cond, "uncond": uncond, "cond_scale": p.cfg_scale, "y_cond": y_cond, "y_uncond": y_uncond}
        sigmas = sigmas.to(x.device)
        fn = getattr(K, K.SAMPLER_TABLE[p.sampler][0])
        if fn in (K.sample_dpmpp_2m, K.sample_lms):
            return fn(cfg, x, sigmas, extra_args=extra)
        if fn in (K.sample_euler_ancestral, K.sample_dpm_2_ancestral, K.sample_dpmpp_2s_ancestral):
            return fn(cfg, x, sigmas, extra_args=extra, eta=p.eta, s_noise=p.s_noise, noise_sampler=rng.next)
        return fn(cfg, x, sigmas, extra_args=extra, s_noise=p.s_noise, noise_sampler=rng.next)  # Euler / Heun / DPM2 (s_churn 0)

    @torch.no_grad()
    def sample(self, p: SamplingParams, cond, uncond, y_cond=None, y_uncond=None):
        """-> final latent [B,4,h,w] fp32 (what `post_sample` hooks see, modules/scripts.py:250)."""
        B = len(p.seeds)
        shape = (4, p.height // 8, p.width // 8)
        rng = ImageRNG(shape, p.seeds, source=p.randn_source, device=self.device)
        x = rng.next()
        sigmas = K.webui_sigmas(self.model_wrap, p.sampler, p.steps)
        x = x * sigmas[0]
        samples = self._run_sampler(p, x, sigmas, rng, cond, uncond, y_cond, y_uncond)
        if not p.enable_hr:
            return samples
        # hires second pass: latent upscale ("Latent" = bilinear) then img2img
        th, tw = int(p.height * p.hr_scale) // 8, int(p.width * p.hr_scale) // 8
        samples = F.interpolate(samples, size=(th, tw), mode="bilinear", antialias=False)
        rng2 = ImageRNG((4, th, tw), p.seeds, source=p.randn_source, device=self.device)
        noise = rng2.next()
        req = p.hr_second_pass_steps or p.steps
        steps, t_enc = K.setup_img2img_steps(req, p.denoising_strength)
        sig = K.webui_sigmas(self.model_wrap, p.sampler, steps)
        sigma_sched = sig[steps - t_enc - 1:]
        xi = samples + noise * sigma_sched[0]
        return self._run_sampler(p, xi, sigma_sched, rng2, cond, uncond, y_cond, y_uncond)

    # -- img2img: modules/processing.py:1728-1790, latent path ------------------------------------------
    @torch.no_grad()
    def img2img(self, p: SamplingParams, encoder, init_images, cond, uncond, encode_noise=None, latent_mask=None,
                y_cond=None, y_uncond=None):
        """init_images [B,3,H,W] in [0,1]; encoder = vae.AutoencoderKLEncode. -> final latent fp32."""
        from .vae import gaussian_sample

        x = init_images.to(self.device, self.dtype_vae) * 2 - 1
        lat = torch.cat([gaussian_sample(encoder.encode_moments(x[i:i + 1]),
                                         None if encode_noise is None else encode_noise[i:i + 1]) for i in range(x.shape[0])])
        init_latent = p.scale_factor * lat.float()
        shape = (4, p.height // 8, p.width // 8)
        rng = ImageRNG(shape, p.seeds, source=p.randn_source, device=self.device)
        noise = rng.next()
        steps, t_enc = K.setup_img2img_steps(p.steps, p.denoising_strength, steps_given=False)
        sig = K.webui_sigmas(self.model_wrap, p.sampler, steps)
        sigma_sched = sig[steps - t_enc - 1:]
        xi = init_latent + noise * sigma_sched[0]
        if latent_mask is None:
            return self._run_sampler(p, xi, sigma_sched, rng, cond, uncond, y_cond, y_uncond), init_latent
        nmask = torch.round(latent_mask.to(self.device, torch.float32)).expand(init_latent.shape)
        mask = 1.0 - nmask
        cfg = CFGDenoiser(self.model_wrap)
        cfg.init_latent, cfg.mask, cfg.nmask = init_latent, mask, nmask
        extra = {"cond": cond, "uncond": uncond, "cond_scale": p.cfg_scale, "y_cond": y_cond, "y_uncond": y_uncond}
        if p.sampler == "Euler a":
            out = K.sample_euler_ancestral(cfg, xi, sigma_sched.to(xi.device), extra_args=extra, eta=p.eta, s_noise=p.s_noise,
                                           noise_sampler=rng.next)
        else:
            out = K.sample_dpmpp_2m(cfg, xi, sigma_sched.to(xi.device), extra_args=extra)
        return out * nmask + init_latent * mask, init_latent

    # -- decode --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, latents, scale_factor):
        """decode_latent_batch: per image, dtype_vae, autocast disabled -> [B,3,H,W] in [0,1] fp32."""
        outs = []
        for i in range(latents.shape[0]):
            z = latents[i:i + 1].to(self.dtype_vae) / scale_factor
            outs.append(self.vae.decode(z)[0])
        x = torch.stack(outs).float()
        return torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0)

    def txt2img(self, p: SamplingParams, cond, uncond, y_cond=None, y_uncond=None):
        lat = self.sample(p, cond, uncond, y_cond, y_uncond)
        return lat, self.decode(lat, p.scale_factor)


def psnr_uint8(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR (dB) on uint8 RGB as the webui would save it: `255. * x` then `astype(np.uint8)`, which truncates
    (modules/processing.py:1034-1035)."""
    a8 = (a.float() * 255.0).clamp(0, 255).floor()
    b8 = (b.float() * 255.0).clamp(0, 255).floor()
    mse = ((a8 - b8) ** 2).mean().item()
    if mse == 0:
        return float("inf")
    import math

    return 10.0 * math.log10(255.0 ** 2 / mse)
