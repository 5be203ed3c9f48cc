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


def slerp(val, low, high):
    """modules/rng.py:85-96, verbatim semantics — including its quirks on [C,h,w] inputs (norms and dot products run
    over dim 1, and nearly parallel inputs fall back to `low * val + high * (1 - val)`)."""
    low_norm = low / torch.norm(low, dim=1, keepdim=True)
    high_norm = high / torch.norm(high, dim=1, keepdim=True)
    dot = (low_norm * high_norm).sum(1)
    if dot.mean() > 0.9995:
        return low * val + high * (1 - val)
    omega = torch.acos(dot)
    so = torch.sin(omega)
    return (torch.sin((1.0 - val) * omega) / so).unsqueeze(1) * low + (torch.sin(val * omega) / so).unsqueeze(1) * high


class ImageRNG:
    """modules/rng.py:99-163. `first()` = the initial latent noise (with the reference's subseed slerp, seed-resize paste
    and eta-noise-seed-delta), `next()` = per-step ancestral noise; both stack per-image draws.

    The reference's `randn(seed, shape)` seeds the GLOBAL generator and draws from it; here a fresh generator seeded with
    `seed` produces the same numbers, and (source != "NV") the global torch seed is set as well so that code running
    after the webui's ImageRNG sees the same global state (`mirror_global_seed`)."""

    def __init__(self, shape, seeds, subseeds=None, subseed_strength=0.0, seed_resize_from_h=0, seed_resize_from_w=0,
                 source: str = "GPU", device="cuda:0", eta_noise_seed_delta: int = 0, mirror_global_seed: bool = True):
        self.shape = tuple(int(s) for s in shape)
        self.seeds = [int(s) for s in seeds]
        self.subseeds = None if subseeds is None else [int(s) for s in subseeds]
        self.subseed_strength = subseed_strength
        self.seed_resize_from_h = seed_resize_from_h
        self.seed_resize_from_w = seed_resize_from_w
        self.eta_noise_seed_delta = int(eta_noise_seed_delta or 0)
        self.mirror_global_seed = mirror_global_seed
        self.source = source
        self.device = torch.device(device)
        self.generators = [self._create_generator(s) for s in self.seeds]
        self.is_first = True

    # modules/rng.py:75-82
    def _create_generator(self, seed):
        if self.source == "NV":
            return PhiloxGenerator(seed)
        gdev = self.device if self.source == "GPU" else torch.device("cpu")
        return torch.Generator(gdev).manual_seed(int(seed))

    def _draw(self, g, shape=None) -> torch.Tensor:
        shape = self.shape if shape is None else shape
        if self.source == "NV":
            return torch.from_numpy(g.randn(shape)).to(self.device, non_blocking=True)
        gdev = self.device if self.source == "GPU" else torch.device("cpu")
        return torch.randn(shape, device=gdev, generator=g).to(self.device)

    # modules/rng.py:6-19: manual_seed(seed), then draw from `generator` or from the freshly seeded global one
    def _randn(self, seed, shape, generator=None) -> torch.Tensor:
        if self.source != "NV" and self.mirror_global_seed:
            torch.manual_seed(int(seed))
        return self._draw(generator if generator is not None else self._create_generator(seed), shape)

    def first(self) -> torch.Tensor:
        resize = self.seed_resize_from_h > 0 and self.seed_resize_from_w > 0
        noise_shape = (self.shape[0], int(self.seed_resize_from_h) // 8, int(self.seed_resize_from_w // 8)) if resize else self.shape
        xs = []
        for i, (seed, generator) in enumerate(zip(self.seeds, self.generators)):
            subnoise = None
            if self.subseeds is not None and self.subseed_strength != 0:
                subseed = 0 if i >= len(self.subseeds) else self.subseeds[i]
                subnoise = self._randn(subseed, noise_shape)
            if noise_shape != self.shape:
                noise = self._randn(seed, noise_shape)
            else:
                noise = self._randn(seed, self.shape, generator=generator)
            if subnoise is not None:
                noise = slerp(self.subseed_strength, noise, subnoise)
            if noise_shape != self.shape:  # paste the centre of the source-resolution noise into fresh target noise
                x = self._randn(seed, self.shape, generator=generator)
                dx = (self.shape[2] - noise_shape[2]) // 2
                dy = (self.shape[1] - noise_shape[1]) // 2
                w = noise_shape[2] if dx >= 0 else noise_shape[2] + 2 * dx
                h = noise_shape[1] if dy >= 0 else noise_shape[1] + 2 * dy
                tx = 0 if dx < 0 else dx
                ty = 0 if dy < 0 else dy
                dx = max(-dx, 0)
                dy = max(-dy, 0)
                x[:, ty:ty + h, tx:tx + w] = noise[:, dy:dy + h, dx:dx + w]
                noise = x
            xs.append(noise)
        if self.eta_noise_seed_delta:
            self.generators = [self._create_generator(seed + self.eta_noise_seed_delta) for seed in self.seeds]
        return torch.stack(xs).to(self.device)

    def next(self) -> torch.Tensor:
        if self.is_first:
            self.is_first = False
            return self.first()
        return torch.stack([self._draw(g) for g in self.generators]).to(self.device)
