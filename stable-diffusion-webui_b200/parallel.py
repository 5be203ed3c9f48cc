"""Multi-GPU: images are independent units (per-sample norms/attention, per-image RNG: modules/rng.py:108,153-163), so
a batch shards one image block per rank with NO per-step collective. The only collective is one NCCL broadcast of
the packed weight blob at load (rank 0 repacks, everyone else receives), over NVLink 5 / NVSwitch.
One process per GPU (torchrun); `gloo` covers the same host logic in CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* come from torchrun. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition of image indices: rank r gets [r*n/world, (r+1)*n/world)."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return list(range(lo, hi))


def shard(seq: Sequence, rank: int, world: int):
    idx = shard_indices(len(seq), rank, world)
    if isinstance(seq, torch.Tensor):
        return seq[idx[0]:idx[-1] + 1] if idx else seq[:0]
    return [seq[i] for i in idx]


def broadcast_weight_blob(blob: torch.Tensor, src: int = 0):
    """One collective for the whole model: every rank has allocated an identically laid-out blob (same config =>
    same packing), rank `src` holds the real contents."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def gather_images(local: torch.Tensor, world: int):
    """Host-side gather of the uint8 results (a few MB per image); rank 0 receives the list, others None."""
    if not dist.is_initialized() or world == 1:
        return [local]
    out = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local.cpu(), out, dst=0)
    return out
