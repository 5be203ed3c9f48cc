"""ORACLE (test infrastructure) — restatement of the webui's noise sources.

modules/rng_philox.py:32-102 (Philox-4x32-10 + Box–Muller reproducing torch.randn(device='cuda') on the CPU; this is
`randn_source="NV"`), and modules/rng.py:99-163 (ImageRNG: one generator per image so that a batch equals the singles).
Pinned by the known-answer vector of modules/rng_philox.py:5-15 and against the reference module itself
(tests/golden/philox_ref.npz, generated from /root/reference by tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def _mulhilo(a: np.ndarray, m: np.uint64):
    p = a.astype(np.uint64) * m
    return (p >> np.uint64(32)).astype(np.uint32), (p & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter [4, n] uint32, key [2, n] uint32 -> [4, n] uint32 (10 rounds; key bumped between rounds)."""
    c0, c1, c2, c3 = (counter[i].copy() for i in range(4))
    k0, k1 = key[0].copy(), key[1].copy()
    for r in range(10):
        hi0, lo0 = _mulhilo(c0, _M0)
        hi1, lo1 = _mulhilo(c2, _M1)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        if r != 9:
            k0 = k0 + _W0
            k1 = k1 + _W1
    return np.stack([c0, c1, c2, c3])


class PhiloxGenerator:
    """rng_philox.Generator: randn(shape) consumes one 'offset' per call; element i uses counter (offset,0,i,0)."""

    def __init__(self, seed: int):
        self.seed = int(seed)
        self.offset = 0

    def randn(self, shape) -> np.ndarray:
        n = int(np.prod(shape))
        counter = np.zeros((4, n), dtype=np.uint32)
        counter[0] = np.uint32(self.offset)
        counter[2] = np.arange(n, dtype=np.uint32)
        self.offset += 1
        seed = np.uint64(self.seed & 0xFFFFFFFFFFFFFFFF)
        key = np.empty((2, n), dtype=np.uint32)
        key[0] = np.uint32(seed & np.uint64(0xFFFFFFFF))
        key[1] = np.uint32(seed >> np.uint64(32))
        with np.errstate(over="ignore"):
            g = philox4x32_10(counter, key)
        two_pow32_inv = np.array([2.3283064e-10], dtype=np.float32)
        two_pow32_inv_2pi = np.array([2.3283064e-10 * 6.2831855], dtype=np.float32)
        u = g[0] * two_pow32_inv + two_pow32_inv / 2
        v = g[1] * two_pow32_inv_2pi + two_pow32_inv_2pi / 2
        s = np.sqrt(-2.0 * np.log(u))
        return (s * np.sin(v)).astype(np.float32).reshape(shape)


class ImageRNG:
    """modules/rng.py:99-163 for the default options (no subseeds, no seed-resize, eta_noise_seed_delta 0).

    source "NV": Philox on the CPU (bit-reproducible anywhere); "GPU": torch.Generator(device) per image
    (the webui default, `randn_source="GPU"`); "CPU": torch CPU generators."""

    def __init__(self, shape, seeds, source="NV", device="cpu"):
        self.shape = tuple(int(s) for s in shape)
        self.seeds = list(seeds)
        self.source = source
        self.device = torch.device(device)
        if source == "NV":
            self.generators = [PhiloxGenerator(s) for s in self.seeds]
        else:
            gdev = self.device if source == "GPU" else torch.device("cpu")
            self.generators = [torch.Generator(gdev).manual_seed(int(s)) for s in self.seeds]

    def _draw(self, g):
        if self.source == "NV":
            return torch.from_numpy(g.randn(self.shape)).to(self.device)
        gdev = self.device if self.source == "GPU" else torch.device("cpu")
        return torch.randn(self.shape, device=gdev, generator=g).to(self.device)

    def next(self) -> torch.Tensor:
        return torch.stack([self._draw(g) for g in self.generators])

    first = next
