"""ORACLE (test infrastructure) — restatement of the KL-VAE (`AutoencoderKL`) decoder / encoder used by SD1.x and SDXL.

Upstream (un-vendored): ldm/modules/diffusionmodules/model.py (Decoder, Encoder, ResnetBlock, AttnBlock, Upsample,
Downsample), ldm/models/autoencoder.py (AutoencoderKL.decode = decoder(post_quant_conv(z))).
In-tree twin of the same arithmetic: modules/models/sd3/sd3_impls.py:171-355 (no post_quant_conv, z=16 there);
AttnBlock.forward as patched by modules/sd_hijack_optimizations.py:637-655 (sdp variant);
config configs/v1-inference.yaml:46-65 (ch 128, ch_mult [1,2,4,4], num_res_blocks 2, z_channels 4, embed_dim 4).
Attribute names reproduce the `first_stage_model.*` state-dict keys.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 4
    embed_dim: int = 4
    in_channels: int = 3


def tiny_vae_config() -> VAEConfig:
    return VAEConfig(ch=64, ch_mult=[1, 2], num_res_blocks=1)


def Normalize(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = Normalize(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = Normalize(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)
        self.cin, self.cout = cin, cout

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.cin != self.cout:
            x = self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    """sdp_attnblock_forward, modules/sd_hijack_optimizations.py:637-655."""

    def __init__(self, c):
        super().__init__()
        self.norm = Normalize(c)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x):
        h_ = self.norm(x)
        q, k, v = self.q(h_), self.k(h_), self.v(h_)
        b, c, h, w = q.shape
        q, k, v = (t.reshape(b, c, h * w).transpose(1, 2).contiguous() for t in (q, k, v))
        out = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        out = out.transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj_out(out)


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Downsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        nres = len(cfg.ch_mult)
        block_in = cfg.ch * cfg.ch_mult[-1]
        self.conv_in = nn.Conv2d(cfg.z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(nres)):
            block = nn.ModuleList()
            block_out = cfg.ch * cfg.ch_mult[i_level]
            for _ in range(cfg.num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            up = nn.Module()
            up.block = block
            if i_level != 0:
                up.upsample = Upsample(block_in)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, cfg.out_ch, 3, padding=1)
        self.nres = nres
        self.num_res_blocks = cfg.num_res_blocks

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        for i_level in reversed(range(self.nres)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        return self.conv_out(F.silu(self.norm_out(h)))


class Encoder(nn.Module):
    """For the img2img / hires non-latent path; mirrors sd3_impls.py:250-302."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        nres = len(cfg.ch_mult)
        self.conv_in = nn.Conv2d(cfg.in_channels, cfg.ch, 3, padding=1)
        in_mult = (1,) + tuple(cfg.ch_mult)
        self.down = nn.ModuleList()
        block_in = cfg.ch
        for i_level in range(nres):
            block = nn.ModuleList()
            block_in = cfg.ch * in_mult[i_level]
            block_out = cfg.ch * cfg.ch_mult[i_level]
            for _ in range(cfg.num_res_blocks):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            down = nn.Module()
            down.block = block
            if i_level != nres - 1:
                down.downsample = Downsample(block_in)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * cfg.z_channels, 3, padding=1)
        self.nres = nres
        self.num_res_blocks = cfg.num_res_blocks

    def forward(self, x):
        h = self.conv_in(x)
        for i_level in range(self.nres):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](h)
            if i_level != self.nres - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        return self.conv_out(F.silu(self.norm_out(h)))


class AutoencoderKLDecode(nn.Module):
    """`first_stage_model` restricted to what decode() touches: post_quant_conv + decoder."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.decoder = Decoder(cfg)
        self.post_quant_conv = nn.Conv2d(cfg.embed_dim, cfg.z_channels, 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    forward = decode


class AutoencoderKLEncode(nn.Module):
    """The other half of `first_stage_model`: encoder + quant_conv -> moments (mean | logvar); ldm AutoencoderKL.encode
    returns DiagonalGaussianDistribution(moments). SURVEY §8(f) N1 (img2img init, modules/sd_samplers_common.py:87-112)."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.z_channels, 2 * cfg.embed_dim, 1)

    def encode_moments(self, x):
        return self.quant_conv(self.encoder(x))

    forward = encode_moments


def gaussian_sample(moments, noise=None):
    """ldm DiagonalGaussianDistribution: logvar clamped to [-30, 20]; sample = mean + std * noise; mode = mean."""
    import torch

    mean, logvar = torch.chunk(moments.float(), 2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise.float()

