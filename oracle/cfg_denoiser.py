"""ORACLE (test infrastructure) — restatement of the per-step CFG inner loop.

modules/sd_samplers_cfg_denoiser.py:156-311 (CFGDenoiser.forward): AND-composed weighted conds (conds_list), the
sub-batched path for unequal cond / uncond token counts (:256-271), s_min_uncond skipping (:218-227,272-275),
batch_cond_uncond=True, no edit model, optional latent mask blend (:174-183, applied after denoising);
combine_denoised follows :74-82 and is pinned to the reference's own function (tests/test_oracle_pins_cpu.py). The dtype plumbing of apply_model follows
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


def stack_conds(tensors):
    """modules/prompt_parser.py:306-317."""
    tensors = list(tensors)
    token_count = max(x.shape[0] for x in tensors)
    for i in range(len(tensors)):
        if tensors[i].shape[0] != token_count:
            tensors[i] = torch.vstack([tensors[i], tensors[i][-1:].repeat([token_count - tensors[i].shape[0], 1])])
    return torch.stack(tensors)


class CFGDenoiser:
    """modules/sd_samplers_cfg_denoiser.py:156-311 with default options (batch_cond_uncond=True, no padding options,
    no edit model). `cond`: [B,T,C] tensor (one cond per image) OR a list (per image) of [(cond [T,C], weight), ...]
    (what reconstruct_multicond_batch flattens, modules/prompt_parser.py:321-349). `uncond`: [B,Tu,C]."""

    def __init__(self, inner_model, mask=None, nmask=None, init_latent=None):
        self.inner_model = inner_model  # CompVisDenoiser
        self.mask, self.nmask, self.init_latent = mask, nmask, init_latent
        self.step = 0

    def __call__(self, x, sigma, uncond, cond, cond_scale, s_min_uncond=0.0, image_cond=None, y_cond=None, y_uncond=None):
        """y_*: SDXL 'vector' conditioning (one row per cond row / per uncond row)."""
        if isinstance(cond, torch.Tensor):
            conds_list = [[(i, 1.0)] for i in range(cond.shape[0])]
            tensor = cond
        else:
            conds_list, rows = [], []
            for per_image in cond:
                mine = []
                for c, w in per_image:
                    mine.append((len(rows), w))
                    rows.append(c)
                conds_list.append(mine)
            tensor = stack_conds(rows)
        batch_size = len(conds_list)
        repeats = [len(conds_list[i]) for i in range(batch_size)]
        x_in = torch.cat([torch.stack([x[i] for _ in range(n)]) for i, n in enumerate(repeats)] + [x])              # :203
        sigma_in = torch.cat([torch.stack([sigma[i] for _ in range(n)]) for i, n in enumerate(repeats)] + [sigma])  # :204
        skip_uncond = bool(self.step % 2 and s_min_uncond > 0 and sigma[0] < s_min_uncond)                          # :218
        if skip_uncond:
            x_in, sigma_in = x_in[:-batch_size], sigma_in[:-batch_size]
        if tensor.shape[1] == uncond.shape[1] or skip_uncond:                                                        # :236
            cond_in = tensor if skip_uncond else torch.cat([tensor, uncond])
            kwargs = {"context": cond_in}
            if y_cond is not None:
                kwargs["y"] = y_cond if skip_uncond else torch.cat([y_cond, y_uncond])
            x_out = self.inner_model(x_in, sigma_in, **kwargs)
        else:                                                                                                        # :256-271
            x_out = torch.zeros_like(x_in)
            sub = batch_size * 2
            for a in range(0, tensor.shape[0], sub):
                b = min(a + sub, tensor.shape[0])
                kwargs = {"context": tensor[a:b]}
                if y_cond is not None:
                    kwargs["y"] = y_cond[a:b]
                x_out[a:b] = self.inner_model(x_in[a:b], sigma_in[a:b], **kwargs)
            kwargs = {"context": uncond}
            if y_uncond is not None:
                kwargs["y"] = y_uncond
            x_out[-uncond.shape[0]:] = self.inner_model(x_in[-uncond.shape[0]:], sigma_in[-uncond.shape[0]:], **kwargs)
        if skip_uncond:                                                                                              # :272-275
            fake_uncond = torch.cat([x_out[c[0][0]:c[0][0] + 1] for c in conds_list])
            x_out = torch.cat([x_out, fake_uncond])
        denoised = combine_denoised(x_out, conds_list, uncond.shape[0], 1.0 if skip_uncond else cond_scale)
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
