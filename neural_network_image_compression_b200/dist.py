"""Multi-GPU plumbing: one process per GPU, images sharded by batch, one allreduce of the symbol histograms.

The codec has no cross-image operation: every rank encodes / decodes its own contiguous
slice of the batch with its own native handle, and the only exchange is a sum-allreduce of the
[3,256] uint64 symbol counts so that every rank holds the same global rate estimate.  The reference has
no distributed code; this is the build's own requirement (BASELINE.json north_star).
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous slice [lo, hi) of `n_items` owned by `rank`; the first n_items % world_size ranks
    hold one extra item.  Slices are disjoint, ordered by rank and cover [0, n_items)."""
    if world_size <= 0 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_process_group_from_env(backend: str | None = None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT / LOCAL_RANK)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_histogram(hist_global, group=None):
    """In-place sum over ranks of the [3,256] symbol counts (torch int64 tensor, CPU for gloo or CUDA
    for NCCL; a NumPy uint64/int64 array is reduced through a CPU tensor).  No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hist_global
    if isinstance(hist_global, np.ndarray):
        t = torch.from_numpy(hist_global.astype(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        hist_global[...] = t.numpy().astype(hist_global.dtype)
        return hist_global
    if hist_global.dtype != torch.int64:
        raise TypeError("histogram tensor must be int64")
    dist.all_reduce(hist_global, op=dist.ReduceOp.SUM, group=group)
    return hist_global


def global_rate(handle, hist_global, lh: int, lw: int, H: int, W: int):
    """Entropy per colour plane (bits/symbol) and bits per pixel of the whole (all-rank) data set from the
    reduced counts: same formula as tf1_13/src/training.py:69-70 with p = global count / global total."""
    from .rate import entropy_from_counts
    ent = entropy_from_counts(handle, hist_global)
    e = ent.detach().cpu().numpy() if hasattr(ent, "detach") else np.asarray(ent)
    bpp = float(np.float32(e.astype(np.float32).sum() * np.float32(lh * lw * 32) / np.float32(H * W)))
    return e, bpp
