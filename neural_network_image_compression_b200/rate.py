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


def _as_host(x):
    """A CPU torch tensor is host memory: hand it to the host path through its NumPy view (shared storage, so results
    written into it are visible to the caller).  CUDA tensors and NumPy arrays pass through."""
    if _is_torch(x) and not x.is_cuda:
        if not x.is_contiguous():
            raise ValueError("CPU tensors passed as buffers must be contiguous")
        return x.numpy()
    return x


def check_count_table(table, shape, on_device: bool, device: int, name: str = "hist_global"):
    """The C side reads and writes prod(shape) 64-bit counters at this address: anything else (a 32-bit table, a float
    tensor, a strided view, a buffer on the wrong side of the bus) would be overrun or silently corrupted, so it is
    rejected here.  Returns the table in the form the call takes (NumPy view for CPU tensors)."""
    table = _as_host(table)
    if _is_torch(table):
        import torch
        if not on_device:
            raise ValueError(f"{name} is a CUDA tensor but the data is in host memory")
        if table.dtype != torch.int64 or tuple(table.shape) != tuple(shape) or not table.is_contiguous():
            raise ValueError(f"{name} must be a contiguous int64 tensor of shape {tuple(shape)}")
        if table.device.index != device:
            raise ValueError(f"{name} must live on the handle's GPU (cuda:{device})")
        return table
    if not isinstance(table, np.ndarray):
        raise TypeError(f"{name} must be a NumPy array or a torch tensor")
    if on_device:
        raise ValueError(f"{name} is in host memory but the data is on the GPU")
    if table.dtype not in (np.uint64, np.int64) or table.shape != tuple(shape) or not table.flags.c_contiguous or not table.flags.writeable:
        raise ValueError(f"{name} must be a writable C-contiguous uint64/int64 array of shape {tuple(shape)}")
    return table


def _check_latent(handle: Handle, latent):
    latent = _as_host(latent)
    if _is_torch(latent):
        import torch
        if latent.dtype != torch.uint8 or latent.dim() != 4 or latent.shape[3] != 96:
            raise ValueError("latent must be a uint8 tensor [N,h,w,96]")
        if latent.device.index != handle.device:
            raise ValueError(f"latent must live on the handle's GPU (cuda:{handle.device})")
        return latent.contiguous(), True
    latent = np.asarray(latent)
    if latent.dtype != np.uint8 or latent.ndim != 4 or latent.shape[3] != 96:
        raise ValueError("latent must be a uint8 array [N,h,w,96]")
    return np.ascontiguousarray(latent), False


def rate(handle: Handle, latent, H: int | None = None, W: int | None = None, hist_global=None) -> Rate:
    """latent: uint8 [N,h,w,96] (NumPy, CPU tensor, or CUDA torch tensor).  H, W default to 8h, 8w.
    `hist_global` (uint64/int64 [3,256], same side of the bus as the latent) is accumulated into when given, so
    micro-batches add up."""
    lib, h = handle.lib, handle.h
    latent, on_device = _check_latent(handle, latent)
    n, lh, lw, _c = latent.shape
    H = 8 * lh if H is None else int(H)
    W = 8 * lw if W is None else int(W)
    if hist_global is not None:
        hist_global = check_count_table(hist_global, (3, 256), on_device, handle.device)
    if on_device:
        import torch
        dev = latent.device
        hist = torch.empty((n, 3, 256), dtype=torch.int32, device=dev)
        ent = torch.empty((n, 3), dtype=torch.float32, device=dev)
        bpp = torch.empty((n,), dtype=torch.float32, device=dev)
        if hist_global is None:
            hist_global = torch.zeros((3, 256), dtype=torch.int64, device=dev)
        handle.check(lib.nnic_rate(h, _ptr(latent), n, lh, lw, H, W, _ptr(hist), _ptr(ent), _ptr(bpp),
                                   _ptr(hist_global), MEM_DEVICE, _stream_of(latent)), "nnic_rate")
        return Rate(hist, ent, bpp, hist_global)
    hist = np.empty((n, 3, 256), np.uint32)
    ent = np.empty((n, 3), np.float32)
    bpp = np.empty((n,), np.float32)
    if hist_global is None:
        hist_global = np.zeros((3, 256), np.uint64)
    handle.check(lib.nnic_rate(h, _ptr(latent), n, lh, lw, H, W, _ptr(hist), _ptr(ent), _ptr(bpp),
                               _ptr(hist_global), MEM_HOST, None), "nnic_rate")
    return Rate(hist, ent, bpp, hist_global)


def rate_channels(handle: Handle, latent, hist_channels=None):
    """Symbol counts per latent FEATURE CHANNEL, uint64/int64 [96,256], summed over the images of `latent` and
    accumulated into `hist_channels` when given.  Rows 32p .. 32p+31 add up to Rate.hist_global[p]."""
    lib, h = handle.lib, handle.h
    latent, on_device = _check_latent(handle, latent)
    n, lh, lw, _c = latent.shape
    if hist_channels is not None:
        hist_channels = check_count_table(hist_channels, (96, 256), on_device, handle.device, "hist_channels")
    if on_device:
        import torch
        if hist_channels is None:
            hist_channels = torch.zeros((96, 256), dtype=torch.int64, device=latent.device)
        handle.check(lib.nnic_rate_channels(h, _ptr(latent), n, lh, lw, _ptr(hist_channels), MEM_DEVICE, _stream_of(latent)),
                     "nnic_rate_channels")
        return hist_channels
    if hist_channels is None:
        hist_channels = np.zeros((96, 256), np.uint64)
    handle.check(lib.nnic_rate_channels(h, _ptr(latent), n, lh, lw, _ptr(hist_channels), MEM_HOST, None), "nnic_rate_channels")
    return hist_channels


def entropy_from_counts(handle: Handle, counts):
    """counts: uint64/int64 [..., 256] (NumPy, CPU tensor or CUDA tensor) -> float32 [...] bits per symbol."""
    lib, h = handle.lib, handle.h
    counts = _as_host(counts)
    if counts.shape[-1] != 256:
        raise ValueError("counts must have 256 bins in the last dimension")
    rows = int(np.prod(counts.shape[:-1]))
    if _is_torch(counts):
        import torch
        if counts.dtype != torch.int64:
            raise ValueError("counts must be int64")
        if counts.device.index != handle.device:
            raise ValueError(f"counts must live on the handle's GPU (cuda:{handle.device})")
        counts = counts.contiguous()
        out = torch.empty(counts.shape[:-1], dtype=torch.float32, device=counts.device)
        handle.check(lib.nnic_entropy_from_counts(h, _ptr(counts), rows, _ptr(out), MEM_DEVICE, _stream_of(counts)),
                     "nnic_entropy_from_counts")
        return out
    counts = np.asarray(counts)
    if counts.dtype not in (np.uint64, np.int64):
        raise ValueError("counts must be uint64 or int64")
    counts = np.ascontiguousarray(counts)
    out = np.empty(counts.shape[:-1], np.float32)
    handle.check(lib.nnic_entropy_from_counts(h, _ptr(counts), rows, _ptr(out), MEM_HOST, None),
                 "nnic_entropy_from_counts")
    return out
