"""ORACLE (test infrastructure) — restatement of the per-step CFG inner loop.

modules/sd_samplers_cfg_denoiser.py:156-311 (CFGDenoiser.forward) for the hot-path configuration: one cond per image
(no AND composition weights other than 1.0 unless given), equal cond / uncond token counts, batch_cond_uncond=True,
no edit model, s_min_uncond = 0, skip_early_cond = 0, optional latent mask blend (:174-183, applied after
denoising); combine_denoised follows :74-82. The dtype plumbing of apply_model follows
modules/sd_hijack_unet.py:40-54 (cast x, t, cond to dtype_unet; run under autocast).
"""
from __future__ import annotations

import torch


def combine_denoised(x_out, conds_list, uncond_batch, cond_scale):
    """sd_samplers_cfg_denoiser.py:74-82."""
    denoised_uncond = x_out[-uncond_batch:]
    denoised = torch.clone(denoised_uncond)
    for i, conds in enumerate(conds_list):
        for cond_index, weight in conds:
            denoised[i] += (x_out[cond_index] - denoised_uncond[i]) * (weight * cond_scale)
    return denoised


class CFGDenoiser:
    def __init__(self, inner_model, mask=None, nmask=None, init_latent=None):
        self.inner_model = inner_model  # CompVisDenoiser
        self.mask, self.nmask, self.init_latent = mask, nmask, init_latent
        self.step = 0

    def __call__(self, x, sigma, uncond, cond, cond_scale, s_min_uncond=0.0, image_cond=None, y_cond=None, y_uncond=None):
        """cond / uncond: [B, T, C] tensors (already `reconstruct_*_batch`-ed); y_*: SDXL 'vector' conditioning."""
        batch_size = x.shape[0]
        conds_list = [[(i, 1.0)] for i in range(batch_size)]
        repeats = [1] * batch_size
        x_in = torch.cat([torch.stack([x[i] for _ in range(n)]) for i, n in enumerate(repeats)] + [x])
        sigma_in = torch.cat([torch.stack([sigma[i] for _ in range(n)]) for i, n in enumerate(repeats)] + [sigma])
        cond_in = torch.cat([cond, uncond])
        kwargs = {"context": cond_in}
        if y_cond is not None:
            kwargs["y"] = torch.cat([y_cond, y_uncond])
        x_out = self.inner_model(x_in, sigma_in, **kwargs)
        denoised = combine_denoised(x_out, conds_list, uncond.shape[0], cond_scale)
        if self.mask is not None:
            denoised = denoised * self.nmask + self.init_latent * self.mask
        self.step += 1
        return denoised


def make_apply_model(unet, dtype_unet: torch.dtype, autocast: bool):
    """LatentDiffusion.apply_model as patched by modules/sd_hijack_unet.py:40-54 + DiffusionWrapper 'crossattn'
    (modules/models/diffusion/ddpm_edit.py:1417-1437): cast inputs to dtype_unet, run the UNet under autocast."""

    def apply_model(x_noisy, t, context=None, y=None):
        x_noisy = x_noisy.to(dtype_unet)
        t = t.to(dtype_unet)
        context = context.to(dtype_unet)
        if y is not None:
            y = y.to(dtype_unet)
        if autocast and x_noisy.is_cuda:
            with torch.autocast("cuda", dtype=dtype_unet):
                return unet(x_noisy, t, context=context, y=y)
        return unet(x_noisy, t, context=context, y=y)

    return apply_model
