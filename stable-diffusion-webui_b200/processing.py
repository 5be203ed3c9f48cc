"""Host-side mirror of the reference pipeline boundary for the accelerated path (boundary A):

  StableDiffusionProcessingTxt2Img     modules/processing.py:1166-1555  (fields used on the hot path only)
  process_images(p) -> Processed       modules/processing.py:819-1112
  decode_latent_batch                  modules/processing.py:625-672
  SdModel.apply_model                  LatentDiffusion.apply_model as patched by modules/sd_hijack_unet.py:40-54
                                       (+ DiffusionWrapper, modules/models/diffusion/ddpm_edit.py:1417-1437;
                                        SDXL: modules/sd_models_xl.py:37-43)

Text encoders, PNG/infotext, scripts and the UI are outside the accelerated path: conditionings arrive as tensors
(what `p.setup_conds()` leaves in p.c / p.uc), images leave as a uint8 tensor.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import lib as L
from . import samplers as S
from .engine import VAEDecoderEngine
from .rng import ImageRNG
from .sd_unet import SdxeUnet

opt_C, opt_f = 4, 8  # modules/processing.py:39-40


class SdModel:
    """What the hot path needs of `shared.sd_model`: the (replacement) UNet, the VAE decoder, the noise schedule."""

    def __init__(self, unet: SdxeUnet, vae: Optional[VAEDecoderEngine], is_sdxl: bool, dtype_unet=torch.float16,
                 device="cuda:0", scale_factor: Optional[float] = None, vae_encoder=None):
        self.unet = unet
        self.vae = vae
        self.vae_encoder = vae_encoder  # engine.VAEEncoderEngine, only needed by img2img
        self.is_sdxl = is_sdxl
        self.dtype_unet = dtype_unet
        self.dtype_vae = dtype_unet
        self.device = torch.device(device)
        self.scale_factor = scale_factor if scale_factor is not None else (0.13025 if is_sdxl else 0.18215)
        self.alphas_cumprod = S.make_alphas_cumprod().to(self.device)
        self.parameterization = "eps"

    def apply_model(self, x_noisy, t, cond=None, **kwargs):
        """cast to dtype_unet, call the UNet through the SdUnet seam (sd_hijack_unet.py:40-54)."""
        ctx = cond["crossattn"] if isinstance(cond, dict) else cond
        vec = cond.get("vector") if isinstance(cond, dict) else None
        return self.apply_model_scaled(x_noisy.to(self.dtype_unet), t, ctx, vec)

    def apply_model_scaled(self, x_in, t, context, vector=None, context_key: int = 0):
        dt = self.dtype_unet
        kwargs = {}
        if context_key:
            kwargs["context_key"] = context_key
        if vector is not None:
            kwargs["y"] = vector.to(dt)
        return self.unet.forward(x_in, t.to(dt), context.to(dt), **kwargs)

    def decode_first_stage(self, z):
        """z already divided by scale_factor upstream? No: the reference's decode_first_stage divides
        (ddpm_edit.py:726-784: z = 1/scale_factor * z). Same here."""
        if self.vae is None:
            raise L.SdxeError("no VAE decoder engine attached")
        return self.vae.decode((z.to(self.dtype_vae) / self.scale_factor).contiguous())


    # ldm LatentDiffusion.encode_first_stage / get_first_stage_encoding: moments -> scale_factor * sample
    def encode_first_stage(self, x):
        if self.vae_encoder is None:
            raise L.SdxeError("no VAE encoder engine attached (img2img needs one)")
        return self.vae_encoder.encode_moments(x.to(self.dtype_vae).contiguous())

    def get_first_stage_encoding(self, moments, noise=None):
        m = moments.float()
        mean, logvar = torch.chunk(m, 2, dim=1)
        if noise is None:  # reference: DiagonalGaussianDistribution.sample() draws from the GLOBAL torch RNG
            noise = torch.randn(mean.shape, device=mean.device, dtype=torch.float32)
        z = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise.float()
        return self.scale_factor * z


def images_tensor_to_samples(image: torch.Tensor, model: SdModel, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """modules/sd_samplers_common.py:87-112 ("Full" VAE encode method): image [B,3,H,W] in [0,1] -> latent [B,4,H/8,W/8].
    The reference encodes image by image; the engine takes the batch (per-sample norms / attention: same result).
    `noise` pins the posterior sample (the reference's comes from the global RNG — SURVEY N1 calls this a parity hazard)."""
    x = image.to(model.device, dtype=model.dtype_vae) * 2 - 1
    return model.get_first_stage_encoding(model.encode_first_stage(x), noise)


@dataclass
class StableDiffusionProcessingTxt2Img:
    sd_model: SdModel = None
    c: object = None                 # cond:  tensor [B,T,C] or {"crossattn","vector"} (per image)
    uc: object = None                # uncond, same form
    seeds: List[int] = field(default_factory=lambda: [1000])
    sampler_name: str = "Euler a"
    scheduler: str = "Automatic"
    steps: int = 20
    cfg_scale: float = 7.0
    width: int = 512
    height: int = 512
    eta: Optional[float] = None
    s_churn: float = 0.0
    s_tmin: float = 0.0
    s_tmax: float = float("inf")
    s_noise: float = 1.0
    hr_scheduler: Optional[str] = None
    s_min_uncond: float = 0.0
    randn_source: str = "GPU"
    subseeds: Optional[List[int]] = None     # modules/processing.py:949 — variation seeds, slerp-ed in at subseed_strength
    subseed_strength: float = 0.0
    seed_resize_from_h: int = 0
    seed_resize_from_w: int = 0
    eta_noise_seed_delta: int = 0            # opts.eta_noise_seed_delta (modules/rng.py:148-150)
    enable_hr: bool = False
    hr_scale: float = 2.0
    hr_second_pass_steps: int = 0
    denoising_strength: float = 0.75
    do_not_decode: bool = False
    check_for_nans: bool = True      # the reference checks unless --disable-nan-check (modules/devices.py:229-231)
    batch_size: int = 0
    rng: ImageRNG = None
    sampler: S.KDiffusionSampler = None
    is_hr_pass: bool = False

    def __post_init__(self):
        self.batch_size = len(self.seeds)

    def make_rng(self, shape, seeds) -> ImageRNG:
        return ImageRNG(shape, seeds, subseeds=self.subseeds, subseed_strength=self.subseed_strength,
                        seed_resize_from_h=self.seed_resize_from_h, seed_resize_from_w=self.seed_resize_from_w,
                        source=self.randn_source, device=self.sd_model.device, eta_noise_seed_delta=self.eta_noise_seed_delta)

    # modules/processing.py:1307-1362
    def sample(self, conditioning, unconditional_conditioning, seeds):
        self.sampler = S.create_sampler(self.sampler_name, self.sd_model)
        x = self.rng.next()
        samples = self.sampler.sample(self, x, conditioning, unconditional_conditioning)
        if not self.enable_hr:
            return samples
        return self.sample_hr_pass(samples, seeds)

    # modules/processing.py:1364-1463, latent upscale mode "Latent" (bilinear, antialias False: shared.py:54-56)
    def sample_hr_pass(self, samples, seeds):
        self.is_hr_pass = True
        tw, th = int(self.width * self.hr_scale), int(self.height * self.hr_scale)
        samples = torch.nn.functional.interpolate(samples, size=(th // opt_f, tw // opt_f), mode="bilinear", antialias=False)
        shape = (opt_C, th // opt_f, tw // opt_f)
        self.rng = self.make_rng(shape, seeds)                                     # processing.py:1429
        noise = self.rng.next()
        self.sampler = S.create_sampler(self.sampler_name, self.sd_model)
        return self.sampler.sample_img2img(self, samples, noise, self.c, self.uc, steps=self.hr_second_pass_steps or self.steps)


@dataclass
class StableDiffusionProcessingImg2Img(StableDiffusionProcessingTxt2Img):
    """modules/processing.py:1527-1790 restricted to the latent path: init image -> VAE encode -> noise at
    denoising_strength -> sampler.sample_img2img -> (optional latent mask blend). Resize modes, PIL mask
    pre-processing, inpainting-model conditioning and colour correction stay upstream of the path."""
    init_images: torch.Tensor = None     # [B,3,H,W] float in [0,1]
    latent_mask: torch.Tensor = None     # [B or 1, 1 or 4, H/8, W/8] float, 1 = repaint (the reference's `nmask`)
    mask_round: bool = True
    inpainting_fill: int = 1             # 1 original (0 "fill" needs the pixel-space blur), 2 latent noise, 3 latent nothing
    encode_noise: torch.Tensor = None    # pins the VAE posterior sample (None: torch.randn, as the reference)
    init_latent: torch.Tensor = None
    mask: torch.Tensor = None
    nmask: torch.Tensor = None

    def init(self, seeds):
        if self.init_images is None:
            raise L.SdxeError("img2img without init_images")
        img = self.init_images
        if img.shape[-2] != self.height or img.shape[-1] != self.width:
            raise L.SdxeError("init_images must already have the target height x width (resize modes are upstream)")
        if img.shape[0] == 1 and self.batch_size > 1:
            img = img.expand(self.batch_size, -1, -1, -1)
        self.init_latent = images_tensor_to_samples(img, self.sd_model, self.encode_noise)
        if self.latent_mask is not None:
            lat = self.latent_mask.to(self.init_latent.device, torch.float32)
            if self.mask_round:
                lat = torch.round(lat)
            lat = lat.expand(self.init_latent.shape)
            self.mask, self.nmask = 1.0 - lat, lat                                  # processing.py:1742-1743
            if self.inpainting_fill == 2:                                           # :1746-1748
                rnd = ImageRNG(tuple(self.init_latent.shape[1:]), seeds, source=self.randn_source, device=self.sd_model.device).next()
                self.init_latent = self.init_latent * self.mask + rnd * self.nmask
            elif self.inpainting_fill == 3:                                         # :1750-1752
                self.init_latent = self.init_latent * self.mask

    # modules/processing.py:1759-1779
    def sample(self, conditioning, unconditional_conditioning, seeds):
        self.init(seeds)
        x = self.rng.next()
        self.sampler = S.create_sampler(self.sampler_name, self.sd_model)
        samples = self.sampler.sample_img2img(self, self.init_latent, x, conditioning, unconditional_conditioning)
        if self.mask is not None:
            samples = samples * self.nmask + self.init_latent * self.mask
        return samples


@dataclass
class Processed:
    images: torch.Tensor = None      # uint8 [B, H, W, 3] on the host (what becomes PIL images in the reference)
    latents: torch.Tensor = None     # final latents fp32 [B,4,h,w] (device)
    seeds: List[int] = None


class NansException(Exception):
    """modules/devices.py:237-238."""


def test_for_nans(x: torch.Tensor, where: str):
    """modules/devices.py:241-265: the probe reads ONE element, `x[(0,) * x.ndim]` (a NaN anywhere in a UNet / VAE
    output spreads to the whole tensor through the next norm or attention), and raises NansException."""
    probe = x[(0,) * x.ndim] if x.ndim else x
    if not bool(torch.isnan(probe)):
        return
    if where == "unet":
        msg = "A tensor with NaNs was produced in Unet."
    elif where == "vae":
        msg = "A tensor with NaNs was produced in VAE."
    else:
        msg = "A tensor with NaNs was produced."
    raise NansException(msg + " Use --disable-nan-check commandline argument to disable this check.")


def decode_latent_batch(model: SdModel, batch: torch.Tensor, target_device=None, check_for_nans=False, batched=True):
    """modules/processing.py:625-672. The reference decodes one image at a time and probes each decoded image
    (:637-641); the engine takes the whole batch in one call (`batched=True`; per-sample GroupNorm / attention make the
    results identical either way) and probes the first element of every image with one host read. The reference's
    NaN -> fp32-VAE retry (:643-665, opt-in via auto_vae_precision) stays upstream of the engine."""
    if batched:
        out = model.decode_first_stage(batch)
    else:
        out = torch.cat([model.decode_first_stage(batch[i:i + 1]) for i in range(batch.shape[0])])
    if check_for_nans and bool(torch.isnan(out[:, 0, 0, 0]).any()):
        test_for_nans(out[int(torch.isnan(out[:, 0, 0, 0]).nonzero()[0])], "vae")
    return out if target_device is None else out.to(target_device)


@torch.no_grad()
def process_images(p: StableDiffusionProcessingTxt2Img, to_host: bool = True) -> Processed:
    """process_images_inner (modules/processing.py:863-1091) for one batch (`n_iter` == 1)."""
    S.state.interrupted = False
    S.state.skipped = False
    dev = p.sd_model.device
    with torch.cuda.device(dev):
        p.rng = p.make_rng((opt_C, p.height // opt_f, p.width // opt_f), p.seeds)           # processing.py:949
        samples = p.sample(p.c, p.uc, p.seeds)
        if p.do_not_decode:
            return Processed(None, samples, list(p.seeds))
        if p.check_for_nans:  # devices.test_for_nans(samples_ddim, "unet"), processing.py:998 (on unless --disable-nan-check)
            test_for_nans(samples, "unet")
        x = decode_latent_batch(p.sd_model, samples, check_for_nans=p.check_for_nans)      # :1002
        x = torch.clamp((x.float() + 1.0) / 2.0, min=0.0, max=1.0)                         # :1004-1005
        # :1034-1035 `x_sample = 255. * ...; x_sample.astype(np.uint8)` — numpy's float -> uint8 cast TRUNCATES
        img = (x.permute(0, 2, 3, 1) * 255.0).to(torch.uint8)
        if not to_host:
            return Processed(img, samples, list(p.seeds))
        host = torch.empty(img.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(img, non_blocking=False)
    return Processed(host, samples, list(p.seeds))
