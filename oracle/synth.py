"""ORACLE (test infrastructure) — seeded synthetic checkpoints and conditioning (no real weights exist offline).

State dicts use the ldm key layout. Initialisation is variance preserving (N(0, 1/fan_in)
scaled per layer kind) so that activations stay O(1) through ~25 blocks and 20-30 sampler steps in 16-bit; the layers
upstream zero-initialises (out_layers.3, proj_out, out.2) get small non-zero weights so tests are not vacuous.
"""
from __future__ import annotations

import math

import torch


@torch.no_grad()
def init_module_(module: torch.nn.Module, seed: int, residual_gain: float = 0.35) -> torch.nn.Module:
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if p.ndim >= 2:
            fan_in = p[0].numel()
            std = 1.0 / math.sqrt(fan_in)
            # damp the last layer of every residual branch: keeps the residual stream from growing with depth
            if any(s in name for s in ("out_layers.3", "proj_out", "to_out.0", "ff.net.2", "conv2.", "attn_1.proj_out")):
                std *= residual_gain
            p.copy_(torch.randn(p.shape, generator=g) * std)
        elif name.endswith("weight"):  # norm scales
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        else:  # biases
            p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return module


def synthetic_context(batch: int, tokens: int, dim: int, seed: int, device="cpu") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, tokens, dim, generator=g).to(device)


def synthetic_vector(batch: int, dim: int, seed: int, device="cpu") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, dim, generator=g).to(device)
