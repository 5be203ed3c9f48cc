"""Noise sources with the webui's semantics (host-side mirror of modules/rng.py:99-163 and modules/rng_philox.py).

`ImageRNG` keeps one generator per image so that an image generated inside a batch — or on another GPU of the shard —
equals the image generated alone (modules/sd_samplers_common.py:206-211 states the requirement).
Sources: "GPU" (torch CUDA generator per image — the webui default), "CPU", "NV" (Philox-4x32-10 + Box–Muller on the
host, bit-identical on every machine).
"""
from __future__ import annotations

import numpy as np
import torch

_PHILOX_M = (np.uint64(0xD2511F53), np.uint64(0xCD9E8D57))
_PHILOX_W = (np.uint32(0x9E3779B9), np.uint32(0xBB67AE85))


def _philox_randn(seed: int, offset: int, n: int) -> np.ndarray:
    """n standard normals for (seed, offset): counter = (offset, 0, i, 0), key = seed; first Box–Muller output only
    (modules/rng_philox.py:84-102)."""
    c0 = np.full(n, offset, dtype=np.uint32)
    c1 = np.zeros(n, dtype=np.uint32)
    c2 = np.arange(n, dtype=np.uint32)
    c3 = np.zeros(n, dtype=np.uint32)
    s = seed & 0xFFFFFFFFFFFFFFFF
    k0 = np.full(n, s & 0xFFFFFFFF, dtype=np.uint32)
    k1 = np.full(n, s >> 32, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = c0.astype(np.uint64) * _PHILOX_M[0]
            p1 = c2.astype(np.uint64) * _PHILOX_M[1]
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r != 9:
                k0 = k0 + _PHILOX_W[0]
                k1 = k1 + _PHILOX_W[1]
    inv = np.array([2.3283064e-10], dtype=np.float32)
    inv2pi = np.array([2.3283064e-10 * 6.2831855], dtype=np.float32)
    u = c0 * inv + inv / 2
    v = c1 * inv2pi + inv2pi / 2
    return (np.sqrt(-2.0 * np.log(u)) * np.sin(v)).astype(np.float32)


class PhiloxGenerator:
    def __init__(self, seed: int):
        self.seed = int(seed)
        self.offset = 0

    def randn(self, shape) -> np.ndarray:
        n = int(np.prod(shape))
        out = _philox_randn(self.seed, self.offset, n).reshape(shape)
        self.offset += 1
        return out


class ImageRNG:
    """`first()` = the initial latent noise, `next()` = per-step ancestral noise; both stack per-image draws.
    Subseed slerp / seed-resize (modules/rng.py:113-146) are not on the benchmarked path and raise if requested."""

    def __init__(self, shape, seeds, subseeds=None, subseed_strength=0.0, seed_resize_from_h=0, seed_resize_from_w=0,
                 source: str = "GPU", device="cuda:0"):
        if (subseeds is not None and subseed_strength != 0) or seed_resize_from_h > 0 or seed_resize_from_w > 0:
            raise NotImplementedError("subseed / seed-resize noise is outside the accelerated path")
        self.shape = tuple(int(s) for s in shape)
        self.seeds = [int(s) for s in seeds]
        self.source = source
        self.device = torch.device(device)
        if source == "NV":
            self.generators = [PhiloxGenerator(s) for s in self.seeds]
        else:
            gdev = self.device if source == "GPU" else torch.device("cpu")
            self.generators = [torch.Generator(gdev).manual_seed(s) for s in self.seeds]
        self.is_first = True

    def _draw(self, g) -> torch.Tensor:
        if self.source == "NV":
            return torch.from_numpy(g.randn(self.shape)).to(self.device, non_blocking=True)
        gdev = self.device if self.source == "GPU" else torch.device("cpu")
        return torch.randn(self.shape, device=gdev, generator=g).to(self.device)

    def first(self) -> torch.Tensor:
        return torch.stack([self._draw(g) for g in self.generators])

    def next(self) -> torch.Tensor:
        if self.is_first:
            self.is_first = False
            return self.first()
        return torch.stack([self._draw(g) for g in self.generators])
