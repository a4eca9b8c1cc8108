"""Mirror of /root/reference/tf2_0/src/decoder.py: Decoder()(x) on the GPU."""
from __future__ import annotations

import numpy as np

from ._lib import MEM_DEVICE, MEM_HOST, _ptr
from .utils import ProClass, _is_torch, _stream_of


class Decoder(ProClass):
    kind = "decoder"

    def __init__(self, device: int = 0, arith: str = "tc_split", handle=None, precision: str = "split"):
        """precision='fp16' selects the optional reduced-precision decoder arithmetic (one fp16 product per MAC;
        reconstructions stay within BASELINE.json's 0.01 dB PSNR of the exact decoder but are not byte-identical)."""
        super().__init__(device, arith, handle)
        if precision != "split":
            self.handle.set_decode_precision(precision)

    def __call__(self, x, return_prequant: bool = False, out=None):
        """decoder.py:39-48.  x: uint8 [N,h,w,96] latent -> uint8 [N,8h,8w,3] RGB (not cropped to the
        source size, like the reference).  NumPy in -> NumPy out through host buffers; CUDA torch
        tensor in -> CUDA tensor out, enqueued on the current stream."""
        lib, h = self.handle.lib, self.handle.h
        if _is_torch(x):
            import torch
            if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 96 or not x.is_cuda:
                raise ValueError("expected a CUDA uint8 tensor [N,h,w,96]")
            if x.device.index != self.device or (out is not None and (not out.is_cuda or out.device != x.device)):
                raise ValueError(f"tensors must live on the handle's GPU (cuda:{self.device})")
            x = x.contiguous()
            n, lh, lw, _ = x.shape
            if out is None:
                out = torch.empty((n, 8 * lh, 8 * lw, 3), dtype=torch.uint8, device=x.device)
            elif tuple(out.shape) != (n, 8 * lh, 8 * lw, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
                raise ValueError("out has the wrong shape, dtype or layout")
            pre = torch.empty(out.shape, dtype=torch.float32, device=x.device) if return_prequant else None
            self.handle.check(lib.nnic_decode(h, _ptr(x), n, lh, lw, _ptr(out), _ptr(pre), MEM_DEVICE,
                                              _stream_of(x)), "nnic_decode")
            return (out, pre) if return_prequant else out
        x = np.asarray(x)
        if x.dtype != np.uint8 or x.ndim != 4 or x.shape[3] != 96:
            raise ValueError("expected a uint8 array [N,h,w,96]")
        x = np.ascontiguousarray(x)
        n, lh, lw, _ = x.shape
        if out is None:
            out = np.empty((n, 8 * lh, 8 * lw, 3), np.uint8)
        elif out.shape != (n, 8 * lh, 8 * lw, 3) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape, dtype or layout")
        pre = np.empty(out.shape, np.float32) if return_prequant else None
        self.handle.check(lib.nnic_decode(h, _ptr(x), n, lh, lw, _ptr(out), _ptr(pre), MEM_HOST, None), "nnic_decode")
        return (out, pre) if return_prequant else out

    def uncompress(self, dataset_path, checkpoint_path=None):
        """decoder.py:50-52: every packed latent PNG of `dataset_path` (a `..._compressed` directory) ->
        reconstruction PNG in the directory named with 'compressed' replaced by 'uncompressed'."""
        return self._use_model(dataset_path, checkpoint_path, dataset_path.replace("compressed", "uncompressed"),
                               in_cshape=96)
