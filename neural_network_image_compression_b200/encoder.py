"""Mirror of /root/reference/tf2_0/src/encoder.py: Encoder()(x) on the GPU."""
from __future__ import annotations

import numpy as np

from ._lib import MEM_DEVICE, MEM_HOST, _ptr
from .utils import ProClass, _is_torch, _stream_of


class Encoder(ProClass):
    kind = "encoder"

    def __call__(self, x, return_prequant: bool = False, out=None):
        """encoder.py:38-47.  x: uint8 [N,H,W,3] RGB -> uint8 [N,ceil(H/8),ceil(W/8),96].

        A NumPy array goes through host buffers (copied to the GPU and back inside the call, like the
        reference's eager tensors).  A CUDA torch.uint8 tensor stays on the device: the result is a
        CUDA tensor and the call only enqueues work on the current stream."""
        lib, h = self.handle.lib, self.handle.h
        if _is_torch(x):
            import torch
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3 or not x.is_cuda:
                raise ValueError("expected a CUDA uint8 tensor [N,H,W,3]")
            if x.device.index != self.device or (out is not None and (not out.is_cuda or out.device != x.device)):
                raise ValueError(f"tensors must live on the handle's GPU (cuda:{self.device})")
            x = x.contiguous()
            n, hh, ww, _ = x.shape
            if out is None:
                out = torch.empty((n, -(-hh // 8), -(-ww // 8), 96), dtype=torch.uint8, device=x.device)
            elif tuple(out.shape) != (n, -(-hh // 8), -(-ww // 8), 96) or out.dtype != torch.uint8 or not out.is_contiguous():
                raise ValueError("out has the wrong shape, dtype or layout")
            pre = torch.empty(out.shape, dtype=torch.float32, device=x.device) if return_prequant else None
            self.handle.check(lib.nnic_encode(h, _ptr(x), n, hh, ww, _ptr(out), _ptr(pre), MEM_DEVICE,
                                              _stream_of(x)), "nnic_encode")
            return (out, pre) if return_prequant else out
        x = np.asarray(x)
        if x.dtype != np.uint8 or x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("expected a uint8 array [N,H,W,3]")
        x = np.ascontiguousarray(x)
        n, hh, ww, _ = x.shape
        if out is None:
            out = np.empty((n, -(-hh // 8), -(-ww // 8), 96), np.uint8)
        elif out.shape != (n, -(-hh // 8), -(-ww // 8), 96) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape, dtype or layout")
        pre = np.empty(out.shape, np.float32) if return_prequant else None
        self.handle.check(lib.nnic_encode(h, _ptr(x), n, hh, ww, _ptr(out), _ptr(pre), MEM_HOST, None), "nnic_encode")
        return (out, pre) if return_prequant else out

    def compress(self, dataset_path, checkpoint_path=None):
        """encoder.py:49-51: every colour image of `dataset_path` -> `<dataset_path>_compressed/<name>.png`, the
        latent packed as a 4h x 8w RGB picture (container.pack_latent)."""
        return self._use_model(dataset_path, checkpoint_path, dataset_path + "_compressed", in_cshape=3)

    def encode_rate(self, x, out=None, hist_global=None):
        """Encoder.__call__ plus the histogram/entropy rate of tf1_13/src/training.py:62-71 in one pass: the
        symbols are counted by the kernel that quantises them (nnic_encode_rate), the latent is not read again.
        Returns (latent, Rate); identical to `lat = enc(x); r = rate(enc.handle, lat, H, W)`."""
        from .rate import Rate, check_count_table
        lib, h = self.handle.lib, self.handle.h
        if _is_torch(x):
            import torch
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3 or not x.is_cuda:
                raise ValueError("expected a CUDA uint8 tensor [N,H,W,3]")
            if x.device.index != self.device or (out is not None and (not out.is_cuda or out.device != x.device)):
                raise ValueError(f"tensors must live on the handle's GPU (cuda:{self.device})")
            x = x.contiguous()
            n, hh, ww, _ = x.shape
            dev = x.device
            if out is None:
                out = torch.empty((n, -(-hh // 8), -(-ww // 8), 96), dtype=torch.uint8, device=dev)
            elif tuple(out.shape) != (n, -(-hh // 8), -(-ww // 8), 96) or out.dtype != torch.uint8 or not out.is_contiguous():
                raise ValueError("out has the wrong shape, dtype or layout")
            hist = torch.empty((n, 3, 256), dtype=torch.int32, device=dev)
            ent = torch.empty((n, 3), dtype=torch.float32, device=dev)
            bpp = torch.empty((n,), dtype=torch.float32, device=dev)
            if hist_global is None:
                hist_global = torch.zeros((3, 256), dtype=torch.int64, device=dev)
            else:
                hist_global = check_count_table(hist_global, (3, 256), True, self.device)
            self.handle.check(lib.nnic_encode_rate(h, _ptr(x), n, hh, ww, _ptr(out), _ptr(hist), _ptr(ent), _ptr(bpp),
                                                   _ptr(hist_global), MEM_DEVICE, _stream_of(x)), "nnic_encode_rate")
            return out, Rate(hist, ent, bpp, hist_global)
        x = np.asarray(x)
        if x.dtype != np.uint8 or x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("expected a uint8 array [N,H,W,3]")
        x = np.ascontiguousarray(x)
        n, hh, ww, _ = x.shape
        if out is None:
            out = np.empty((n, -(-hh // 8), -(-ww // 8), 96), np.uint8)
        elif out.shape != (n, -(-hh // 8), -(-ww // 8), 96) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape, dtype or layout")
        hist = np.empty((n, 3, 256), np.uint32)
        ent = np.empty((n, 3), np.float32)
        bpp = np.empty((n,), np.float32)
        if hist_global is None:
            hist_global = np.zeros((3, 256), np.uint64)
        else:
            hist_global = check_count_table(hist_global, (3, 256), False, self.device)
        self.handle.check(lib.nnic_encode_rate(h, _ptr(x), n, hh, ww, _ptr(out), _ptr(hist), _ptr(ent), _ptr(bpp),
                                               _ptr(hist_global), MEM_HOST, None), "nnic_encode_rate")
        return out, Rate(hist, ent, bpp, hist_global)
