"""Rate estimate of a uint8 latent: the discrete histogram entropy of
/root/reference/tf1_13/src/training.py:62-71 (per image and colour plane), on the GPU."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import MEM_DEVICE, MEM_HOST, Handle, _ptr
from .utils import _is_torch, _stream_of


@dataclass
class Rate:
    hist: object          # uint32 [N,3,256]
    entropy_bits: object  # float32 [N,3]   bits per symbol
    bpp: object           # float32 [N]     sum_p entropy * (h*w*32) / (H*W)
    hist_global: object   # uint64 [3,256]  summed over the N images (plus whatever was passed in)


def rate(handle: Handle, latent, H: int | None = None, W: int | None = None, hist_global=None) -> Rate:
    """latent: uint8 [N,h,w,96] (NumPy or CUDA torch tensor).  H, W default to 8h, 8w.
    `hist_global` (uint64/int64 [3,256]) is accumulated into when given, so micro-batches add up."""
    lib, h = handle.lib, handle.h
    n, lh, lw, c = latent.shape
    if c != 96:
        raise ValueError("latent must have 96 channels")
    H = 8 * lh if H is None else int(H)
    W = 8 * lw if W is None else int(W)
    if _is_torch(latent):
        import torch
        if not latent.is_cuda or latent.device.index != handle.device:
            raise ValueError(f"latent must live on the handle's GPU (cuda:{handle.device})")
        latent = latent.contiguous()
        dev = latent.device
        hist = torch.empty((n, 3, 256), dtype=torch.int32, device=dev)
        ent = torch.empty((n, 3), dtype=torch.float32, device=dev)
        bpp = torch.empty((n,), dtype=torch.float32, device=dev)
        if hist_global is None:
            hist_global = torch.zeros((3, 256), dtype=torch.int64, device=dev)
        handle.check(lib.nnic_rate(h, _ptr(latent), n, lh, lw, H, W, _ptr(hist), _ptr(ent), _ptr(bpp),
                                   _ptr(hist_global), MEM_DEVICE, _stream_of(latent)), "nnic_rate")
        return Rate(hist, ent, bpp, hist_global)
    latent = np.ascontiguousarray(latent, np.uint8)
    hist = np.empty((n, 3, 256), np.uint32)
    ent = np.empty((n, 3), np.float32)
    bpp = np.empty((n,), np.float32)
    if hist_global is None:
        hist_global = np.zeros((3, 256), np.uint64)
    handle.check(lib.nnic_rate(h, _ptr(latent), n, lh, lw, H, W, _ptr(hist), _ptr(ent), _ptr(bpp),
                               _ptr(hist_global), MEM_HOST, None), "nnic_rate")
    return Rate(hist, ent, bpp, hist_global)


def entropy_from_counts(handle: Handle, counts):
    """counts: uint64/int64 [rows,256] (NumPy or CUDA tensor) -> float32 [rows] bits per symbol."""
    lib, h = handle.lib, handle.h
    rows = int(np.prod(counts.shape[:-1]))
    if _is_torch(counts):
        import torch
        counts = counts.contiguous()
        out = torch.empty(counts.shape[:-1], dtype=torch.float32, device=counts.device)
        handle.check(lib.nnic_entropy_from_counts(h, _ptr(counts), rows, _ptr(out), MEM_DEVICE, _stream_of(counts)),
                     "nnic_entropy_from_counts")
        return out
    counts = np.ascontiguousarray(counts, np.uint64)
    out = np.empty(counts.shape[:-1], np.float32)
    handle.check(lib.nnic_entropy_from_counts(h, _ptr(counts), rows, _ptr(out), MEM_HOST, None),
                 "nnic_entropy_from_counts")
    return out
