"""Host-side mirror of /root/reference/tf2_0/src/utils.py for the codec hot path.

Same names and argument meaning as the reference: ProClass holds the pair of networks ('Y', 'CbCr'),
`run_model` is the plane-level call with weight sets (0, 1, 1) (utils.py:19-24) and `load` reads one
weight file per network at `path + 'Y'` / `path + 'CbCr'` (utils.py:26-28).  The arithmetic runs in
libnnic.so on the GPU; nothing here computes on the CPU.
"""
from __future__ import annotations

import os

import numpy as np

from . import weights as W
from ._lib import MEM_DEVICE, MEM_HOST, Handle, NnicError, _ptr
from .container import DatasetDriver, read_dataset, save_img  # noqa: F401  (utils.py:30-62,85-120)

_MODELS = list(W.MODEL_SUFFIXES)


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _stream_of(x):
    import torch
    return torch.cuda.current_stream(x.device).cuda_stream


class ProClass(DatasetDriver):
    """Pair of networks of one kind ('encoder' or 'decoder') on one GPU."""

    kind = None  # set by subclasses

    def __init__(self, device: int = 0, arith: str = "tc_split", handle: Handle | None = None):
        self.handle = handle if handle is not None else Handle(device, arith)
        self.device = self.handle.device
        self._set_base = W.SET_ENC_Y if self.kind == "encoder" else W.SET_DEC_Y
        self.weights = [None, None]

    # ---- weights -----------------------------------------------------------------------------
    def set_weights(self, model_index: int, w: dict):
        """Install one network's weights ({'<layer>/kernel', '<layer>/bias'} in Keras layouts)."""
        W.check_weight_set(self.kind, w)
        for li, (name, *_rest) in enumerate(W.layers_of(self.kind)):
            self.handle.set_weights(self._set_base + model_index, li, w[name + "/kernel"], w[name + "/bias"])
        self.weights[model_index] = w

    def init_random(self, seeds=None, gain: float = 1.0, bias_range: float = 0.0):
        """Keras-default initialisation (what a freshly constructed reference model holds), drawn by the library itself
        (nnic_init_random_scaled: NumPy's default_rng stream, so the networks equal weights.glorot_uniform(kind, seed));
        `self.weights` keeps the same arrays for save() and for the parity tests."""
        prefix = "enc" if self.kind == "encoder" else "dec"
        for i, name in enumerate(_MODELS):
            seed = (seeds or W.DEFAULT_SEEDS)[prefix + name] if not isinstance(seeds, (list, tuple)) else seeds[i]
            self.handle.check(self.handle.lib.nnic_init_random_scaled(self.handle.h, self._set_base + i, int(seed), float(gain),
                                                                      float(bias_range)), "nnic_init_random_scaled")
            self.weights[i] = W.glorot_uniform_native(self.kind, seed, gain, bias_range)
        return self

    def load(self, path: str):
        """utils.py:26-28: one weight file per network, at path+'Y' and path+'CbCr'.

        Accepts what the reference writes there -- a TensorFlow checkpoint `<path><name>.index` +
        `.data-00000-of-00001` (Keras `save_weights`, read by tfbundle.py without TensorFlow) -- or `<path><name>.npz`
        (weights.save_npz) holding the same variables in the same layouts."""
        from . import tfbundle
        names = [layer[0] for layer in W.layers_of(self.kind)]
        for i, name in enumerate(_MODELS):
            p = path + str(name)
            if tfbundle.is_bundle(p):
                self.set_weights(i, tfbundle.keras_weights(p, names))
                continue
            if not p.endswith(".npz"):
                p += ".npz"
            if not os.path.exists(p):
                raise FileNotFoundError(p)
            self.set_weights(i, W.load_npz(p))
        return self

    def save(self, path: str):
        for i, name in enumerate(_MODELS):
            W.save_npz(path + str(name) + ".npz", self.weights[i])

    # ---- plane-level model call ----------------------------------------------------------------
    def run_model(self, x):
        """utils.py:19-24.  x: three float32 arrays [N,H,W,C] (C=1 for the encoder, 32 for the
        decoder); returns three float32 arrays (models[0](x[0]), models[1](x[1]), models[1](x[2]))."""
        planes = np.ascontiguousarray(np.stack([np.asarray(p, np.float32) for p in x], axis=0))
        _, n, hh, ww, c = planes.shape
        lib, h = self.handle.lib, self.handle.h
        if self.kind == "encoder":
            if c != 1:
                raise ValueError("encoder planes must have one channel")
            out = np.empty((3, n, -(-hh // 8), -(-ww // 8), 32), np.float32)
            self.handle.check(lib.nnic_run_encoder_planes(h, _ptr(planes), n, hh, ww, _ptr(out), MEM_HOST, None),
                              "nnic_run_encoder_planes")
        else:
            if c != 32:
                raise ValueError("decoder planes must have 32 channels")
            out = np.empty((3, n, hh * 8, ww * 8, 1), np.float32)
            self.handle.check(lib.nnic_run_decoder_planes(h, _ptr(planes), n, hh, ww, _ptr(out), MEM_HOST, None),
                              "nnic_run_decoder_planes")
        return [out[0], out[1], out[2]]
