"""ORACLE (test infrastructure) — restatement of the ldm / sgm `UNetModel` with the webui's patches applied.

Upstream (un-vendored, pinned at modules/launch_utils.py:355-356): ldm/modules/diffusionmodules/openaimodel.py
(UNetModel, ResBlock, Upsample, Downsample, TimestepEmbedSequential), ldm/modules/attention.py (SpatialTransformer,
BasicTransformerBlock, CrossAttention, FeedForward, GEGLU); sgm has the same lineage for SDXL.
In-tree anchors: state-dict key layout extensions-builtin/Lora/networks.py:43-119; timestep embedding
modules/sd_hijack_unet.py:58-78; SpatialTransformer.forward modules/sd_hijack_unet.py:83-102; attention
modules/sd_hijack_optimizations.py:508-546 (sdp variant); GroupNorm32 fp32 upcast modules/devices.py:284-295;
configs configs/v1-inference.yaml:29-44 and configs/sd_xl_inpaint.yaml:19-37.

Module attribute names reproduce the ldm state-dict keys exactly, so one checkpoint loads into oracle and engine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    channel_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    transformer_depth: List[int] = field(default_factory=lambda: [1, 1, 1, 0])  # per level, 0 = no attention
    num_heads: int = 8              # > 0: fixed head count; else num_head_channels
    num_head_channels: int = -1
    context_dim: int = 768
    use_linear_in_transformer: bool = False
    adm_in_channels: int = 0        # SDXL: 2816
    middle_depth: int = 1           # transformer depth of the middle block (SD1.x: 1, SDXL: 10)

    def heads_for(self, ch: int):
        if self.num_head_channels and self.num_head_channels > 0:
            return ch // self.num_head_channels, self.num_head_channels
        return self.num_heads, ch // self.num_heads


def sd15_config() -> UNetConfig:
    """configs/v1-inference.yaml:29-44 (attention_resolutions [4,2,1], num_heads 8, transformer_depth 1)."""
    return UNetConfig()


def sdxl_config() -> UNetConfig:
    """SDXL base (sgm sd_xl_base.yaml; in-tree twin configs/sd_xl_inpaint.yaml:19-37 with in_channels 4)."""
    return UNetConfig(model_channels=320, channel_mult=[1, 2, 4], transformer_depth=[0, 2, 10], num_heads=-1,
                      num_head_channels=64, context_dim=2048, use_linear_in_transformer=True, adm_in_channels=2816,
                      middle_depth=10)


def tiny_config(linear: bool = False, adm: int = 0) -> UNetConfig:
    """A structurally complete small UNet for fast tests (2 levels, attention at both, 64-channel granularity)."""
    return UNetConfig(model_channels=64, channel_mult=[1, 2], transformer_depth=[1, 2], num_heads=-1 if linear else 2,
                      num_head_channels=64 if linear else -1, context_dim=128, use_linear_in_transformer=linear,
                      adm_in_channels=adm)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """modules/sd_hijack_unet.py:58-78 — cos first, then sin; computed in fp32 on the timesteps' device."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    """ldm util.GroupNorm32: statistics in fp32, result cast back (modules/devices.py:284-295 states the upcast)."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class Upsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class Downsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class ResBlock(nn.Module):
    def __init__(self, ch, emb_ch, out_ch):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(ch, out_ch, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_ch, out_ch))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_ch), nn.SiLU(), nn.Dropout(0.0), nn.Conv2d(out_ch, out_ch, 3, padding=1))
        self.skip_connection = nn.Identity() if ch == out_ch else nn.Conv2d(ch, out_ch, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        emb_out = self.emb_layers(emb).type(h.dtype)
        h = h + emb_out[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class CrossAttention(nn.Module):
    """ldm attention.CrossAttention with forward = scaled_dot_product_attention_forward
    (modules/sd_hijack_optimizations.py:508-546)."""

    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        context_dim = context_dim or query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5  # ldm attribute the non-SDP reference forwards read (sd_hijack_optimizations.py:236)
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        b, n, _ = x.shape
        h = self.heads
        q_in = self.to_q(x)
        context = x if context is None else context
        k_in = self.to_k(context)
        v_in = self.to_v(context)
        d = q_in.shape[-1] // h
        q = q_in.view(b, -1, h, d).transpose(1, 2)
        k = k_in.view(b, -1, h, d).transpose(1, 2)
        v = v_in.view(b, -1, h, d).transpose(1, 2)
        dtype = q.dtype
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        out = out.transpose(1, 2).reshape(b, -1, h * d).to(dtype)
        return self.to_out(out)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class SpatialTransformer(nn.Module):
    """forward follows modules/sd_hijack_unet.py:83-102."""

    def __init__(self, ch, heads, dim_head, depth, context_dim, use_linear):
        super().__init__()
        inner = heads * dim_head
        self.use_linear = use_linear
        self.norm = nn.GroupNorm(32, ch, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(ch, inner) if use_linear else nn.Conv2d(ch, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, context_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(inner, ch) if use_linear else nn.Conv2d(inner, ch, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        x_in = x
        x = self.norm(x)
        if not self.use_linear:
            x = self.proj_in(x)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        if self.use_linear:
            x = self.proj_in(x)
        for block in self.transformer_blocks:
            x = block(x, context=context)
        if self.use_linear:
            x = self.proj_out(x)
        x = x.view(b, h, w, c).permute(0, 3, 1, 2)
        if not self.use_linear:
            x = self.proj_out(x)
        return x + x_in


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        mc = cfg.model_channels
        ted = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        if cfg.adm_in_channels:
            self.label_emb = nn.Sequential(nn.Sequential(nn.Linear(cfg.adm_in_channels, ted), nn.SiLU(), nn.Linear(ted, ted)))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(cfg.in_channels, mc, 3, padding=1))])
        chans = [mc]
        ch = mc
        nl = len(cfg.channel_mult)

        def st(ch, depth):
            heads, dh = cfg.heads_for(ch)
            return SpatialTransformer(ch, heads, dh, depth, cfg.context_dim, cfg.use_linear_in_transformer)

        for level, mult in enumerate(cfg.channel_mult):
            for _ in range(cfg.num_res_blocks):
                layers = [ResBlock(ch, ted, mult * mc)]
                ch = mult * mc
                if cfg.transformer_depth[level] > 0:
                    layers.append(st(ch, cfg.transformer_depth[level]))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != nl - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch)))
                chans.append(ch)
        mid_depth = cfg.middle_depth
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, ch), st(ch, mid_depth), ResBlock(ch, ted, ch))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
            for i in range(cfg.num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, mc * mult)]
                ch = mc * mult
                if cfg.transformer_depth[level] > 0:
                    layers.append(st(ch, cfg.transformer_depth[level]))
                if level and i == cfg.num_res_blocks:
                    layers.append(Upsample(ch))
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(mc, cfg.out_channels, 3, padding=1))

    def forward(self, x, timesteps=None, context=None, y=None):
        hs = []
        t_emb = timestep_embedding(timesteps, self.cfg.model_channels).to(x.dtype)  # cast: sd_hijack_unet.py:145-154
        emb = self.time_embed(t_emb)
        if self.cfg.adm_in_channels:
            emb = emb + self.label_emb(y)
        h = x
        for module in self.input_blocks:
            h = module(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for module in self.output_blocks:
            h = torch.cat([h, hs.pop()], dim=1)
            h = module(h, emb, context)
        h = h.type(x.dtype)
        return self.out(h)
