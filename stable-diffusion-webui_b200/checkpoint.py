"""Checkpoint layout of the hot-path models: every state-dict key and shape of the ldm / sgm UNet and of the KL-VAE
decoder, derived from the architecture spec (key patterns corroborated in-tree by
extensions-builtin/Lora/networks.py:43-119 and modules/sd_hijack_optimizations.py:556-605), plus seeded synthetic
initialisation for benchmarks and tests (no real checkpoints exist offline).

`unet_param_shapes(UNetSpec.sd15())` sums to 859,520,964 parameters, `UNetSpec.sdxl()` to 2,567,463,684 and
`vae_decoder_param_shapes(VAESpec())` to 49,490,199 (decoder 49,490,179 + post_quant_conv 20).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .engine import UNetSpec, VAESpec

Shapes = "OrderedDict[str, Tuple[int, ...]]"


def _lin(d, p, n, k, bias=True):
    d[p + ".weight"] = (n, k)
    if bias:
        d[p + ".bias"] = (n,)


def _conv(d, p, co, ci, k):
    d[p + ".weight"] = (co, ci, k, k)
    d[p + ".bias"] = (co,)


def _norm(d, p, c):
    d[p + ".weight"] = (c,)
    d[p + ".bias"] = (c,)


def _res(d, p, cin, cout, ted):
    _norm(d, p + ".in_layers.0", cin)
    _conv(d, p + ".in_layers.2", cout, cin, 3)
    _lin(d, p + ".emb_layers.1", cout, ted)
    _norm(d, p + ".out_layers.0", cout)
    _conv(d, p + ".out_layers.3", cout, cout, 3)
    if cin != cout:
        _conv(d, p + ".skip_connection", cout, cin, 1)


def _st(d, p, c, depth, ctx, linear):
    _norm(d, p + ".norm", c)
    if linear:
        _lin(d, p + ".proj_in", c, c)
    else:
        _conv(d, p + ".proj_in", c, c, 1)
    for j in range(depth):
        b = f"{p}.transformer_blocks.{j}"
        for a, kd in (("attn1", c), ("attn2", ctx)):
            _lin(d, f"{b}.{a}.to_q", c, c, bias=False)
            _lin(d, f"{b}.{a}.to_k", c, kd, bias=False)
            _lin(d, f"{b}.{a}.to_v", c, kd, bias=False)
            _lin(d, f"{b}.{a}.to_out.0", c, c)
        _lin(d, f"{b}.ff.net.0.proj", 8 * c, c)
        _lin(d, f"{b}.ff.net.2", c, 4 * c)
        for nname in ("norm1", "norm2", "norm3"):
            _norm(d, f"{b}.{nname}", c)
    if linear:
        _lin(d, p + ".proj_out", c, c)
    else:
        _conv(d, p + ".proj_out", c, c, 1)


def unet_param_shapes(spec: UNetSpec) -> "OrderedDict[str, Tuple[int, ...]]":
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    mc, ted = spec.model_channels, 4 * spec.model_channels
    lin = spec.use_linear_in_transformer
    _lin(d, "time_embed.0", ted, mc)
    _lin(d, "time_embed.2", ted, ted)
    if spec.adm_in_channels:
        _lin(d, "label_emb.0.0", ted, spec.adm_in_channels)
        _lin(d, "label_emb.0.2", ted, ted)
    _conv(d, "input_blocks.0.0", mc, spec.in_channels, 3)
    chans, ch, idx = [mc], mc, 1
    nl = len(spec.channel_mult)
    for level, mult in enumerate(spec.channel_mult):
        for _ in range(spec.num_res_blocks):
            _res(d, f"input_blocks.{idx}.0", ch, mult * mc, ted)
            ch = mult * mc
            if spec.transformer_depth[level] > 0:
                _st(d, f"input_blocks.{idx}.1", ch, spec.transformer_depth[level], spec.context_dim, lin)
            chans.append(ch)
            idx += 1
        if level != nl - 1:
            _conv(d, f"input_blocks.{idx}.0.op", ch, ch, 3)
            chans.append(ch)
            idx += 1
    _res(d, "middle_block.0", ch, ch, ted)
    _st(d, "middle_block.1", ch, spec.middle_depth, spec.context_dim, lin)
    _res(d, "middle_block.2", ch, ch, ted)
    idx = 0
    for level in reversed(range(nl)):
        mult = spec.channel_mult[level]
        for i in range(spec.num_res_blocks + 1):
            ich = chans.pop()
            _res(d, f"output_blocks.{idx}.0", ch + ich, mc * mult, ted)
            ch = mc * mult
            sub = 1
            if spec.transformer_depth[level] > 0:
                _st(d, f"output_blocks.{idx}.1", ch, spec.transformer_depth[level], spec.context_dim, lin)
                sub = 2
            if level and i == spec.num_res_blocks:
                _conv(d, f"output_blocks.{idx}.{sub}.conv", ch, ch, 3)
            idx += 1
    _norm(d, "out.0", ch)
    _conv(d, "out.2", spec.out_channels, ch, 3)
    return d


def _vres(d, p, cin, cout):
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, p + ".nin_shortcut", cout, cin, 1)


def vae_decoder_param_shapes(spec: VAESpec, embed_dim: int = 4) -> "OrderedDict[str, Tuple[int, ...]]":
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    nl = len(spec.ch_mult)
    bi = spec.ch * spec.ch_mult[-1]
    _conv(d, "post_quant_conv", spec.z_channels, embed_dim, 1)
    _conv(d, "decoder.conv_in", bi, spec.z_channels, 3)
    _vres(d, "decoder.mid.block_1", bi, bi)
    _norm(d, "decoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        _conv(d, f"decoder.mid.attn_1.{n}", bi, bi, 1)
    _vres(d, "decoder.mid.block_2", bi, bi)
    for level in reversed(range(nl)):
        bo = spec.ch * spec.ch_mult[level]
        for j in range(spec.num_res_blocks + 1):
            _vres(d, f"decoder.up.{level}.block.{j}", bi, bo)
            bi = bo
        if level != 0:
            _conv(d, f"decoder.up.{level}.upsample.conv", bi, bi, 3)
    _norm(d, "decoder.norm_out", bi)
    _conv(d, "decoder.conv_out", spec.out_ch, bi, 3)
    return d


def vae_encoder_param_shapes(spec: VAESpec, embed_dim: int = 4) -> "OrderedDict[str, Tuple[int, ...]]":
    """ldm Encoder + quant_conv (the other half of `first_stage_model`), state-dict order."""
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _conv(d, "encoder.conv_in", spec.ch, spec.out_ch, 3)
    bi = spec.ch
    nl = len(spec.ch_mult)
    for level in range(nl):
        bo = spec.ch * spec.ch_mult[level]
        for j in range(spec.num_res_blocks):
            _vres(d, f"encoder.down.{level}.block.{j}", bi, bo)
            bi = bo
        if level != nl - 1:
            _conv(d, f"encoder.down.{level}.downsample.conv", bi, bi, 3)
    _vres(d, "encoder.mid.block_1", bi, bi)
    _norm(d, "encoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        _conv(d, f"encoder.mid.attn_1.{n}", bi, bi, 1)
    _vres(d, "encoder.mid.block_2", bi, bi)
    _norm(d, "encoder.norm_out", bi)
    _conv(d, "encoder.conv_out", 2 * spec.z_channels, bi, 3)
    _conv(d, "quant_conv", 2 * embed_dim, 2 * spec.z_channels, 1)
    return d


def param_count(shapes) -> int:
    return sum(math.prod(s) for s in shapes.values())


_RESIDUAL_TAILS = ("out_layers.3", "proj_out", "to_out.0", "ff.net.2", "conv2.", "attn_1.proj_out")


@torch.no_grad()
def synthetic_state_dict(shapes, seed: int, device="cpu", dtype=torch.float32, residual_gain: float = 0.35) -> Dict[str, torch.Tensor]:
    """Variance-preserving seeded init (activations stay O(1) through the whole net and 20-30 sampler steps in 16 bit);
    the layers upstream zero-initialises get small non-zero weights so nothing is vacuous."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shape in shapes.items():
        if len(shape) >= 2:
            fan_in = math.prod(shape[1:])
            std = 1.0 / math.sqrt(fan_in)
            if any(s in k for s in _RESIDUAL_TAILS):
                std *= residual_gain
            t = torch.randn(shape, generator=g, device=device) * std
        elif k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            t = 0.05 * torch.randn(shape, generator=g, device=device)
        sd[k] = t.to(dtype)
    return sd


def empty_state_dict(shapes, device, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Uninitialised tensors of the right shapes: what non-root ranks ingest before the weight-blob broadcast."""
    return {k: torch.empty(s, device=device, dtype=dtype) for k, s in shapes.items()}
