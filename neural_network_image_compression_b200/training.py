"""Forward side of /root/reference/tf2_0/src/training.py on the GPU: the computations the reference's
training step runs beside encoder and decoder -- `Entropynet` (training.py:25-42), the uniform-noise quantisation proxy
(training.py:87-88) and the SSIM term of the loss (training.py:108-119).  Forward only: gradients, optimisers and the
training loop itself are outside this build's scope.  Everything runs in libnnic.so; there is no CPU path."""
from __future__ import annotations

import numpy as np

from ._lib import MEM_DEVICE, MEM_HOST, Handle, _ptr
from .utils import _is_torch, _stream_of

# (name, kernel shape) -- training.py:28-33; dense1's input size depends on the latent size (Flatten)
ENTROPYNET_CONVS = (("conv1", (5, 5, 32, 64)), ("conv2", (3, 3, 64, 64)), ("conv3", (3, 3, 64, 64)))


def _f32_in(handle: Handle, x, what: str):
    """float32 input as (buffer, on_device): NumPy / CPU tensor -> host path, CUDA tensor on the handle's GPU -> device path."""
    if _is_torch(x):
        import torch
        if x.dtype != torch.float32:
            raise ValueError(f"{what} must be float32")
        if not x.is_cuda:
            return np.ascontiguousarray(x.numpy()), False
        if x.device.index != handle.device:
            raise ValueError(f"{what} must live on the handle's GPU (cuda:{handle.device})")
        return x.contiguous(), True
    x = np.asarray(x)
    if x.dtype != np.float32:
        raise ValueError(f"{what} must be float32")
    return np.ascontiguousarray(x), False


class Entropynet:
    """training.py:25-42.  `Entropynet()(x)`: x float32 [P,h,w,32] (the encoder's output in [0,1], planes stacked on the
    batch axis) -> float32 [P,1], the network's estimate of the PNG rate in bits per pixel, clipped to [0, 8]."""

    def __init__(self, device: int = 0, arith: str = "tc_split", handle: Handle | None = None):
        self.handle = handle if handle is not None else Handle(device, arith)
        self.device = self.handle.device
        self.weights = None

    def set_weights(self, w: dict):
        """w: {'conv1/kernel', 'conv1/bias', ..., 'dense1/kernel' [F,512], 'dense1/bias', 'dense2/kernel' [512,1], 'dense2/bias'}
        in the Keras layouts."""
        lib, h = self.handle.lib, self.handle.h
        for li, (name, shape) in enumerate(ENTROPYNET_CONVS):
            k, b = np.ascontiguousarray(w[name + "/kernel"], np.float32), np.ascontiguousarray(w[name + "/bias"], np.float32)
            if k.shape != shape or b.shape != (shape[3],):
                raise ValueError(f"{name}: kernel {k.shape} / bias {b.shape}, expected {shape} / {(shape[3],)}")
            self.handle.check(lib.nnic_entropynet_set_weights(h, li, _ptr(k), _ptr(b), 0), "nnic_entropynet_set_weights")
        k, b = np.ascontiguousarray(w["dense1/kernel"], np.float32), np.ascontiguousarray(w["dense1/bias"], np.float32)
        if k.ndim != 2 or k.shape[1] != 512 or k.shape[0] % 64 or b.shape != (512,):
            raise ValueError("dense1/kernel must be [64*h2*w2, 512] with a [512] bias")
        self.handle.check(lib.nnic_entropynet_set_weights(h, 3, _ptr(k), _ptr(b), k.shape[0]), "nnic_entropynet_set_weights")
        k, b = np.ascontiguousarray(w["dense2/kernel"], np.float32), np.ascontiguousarray(w["dense2/bias"], np.float32)
        if k.shape != (512, 1) or b.shape != (1,):
            raise ValueError("dense2/kernel must be [512, 1] with a [1] bias")
        self.handle.check(lib.nnic_entropynet_set_weights(h, 4, _ptr(k), _ptr(b), 0), "nnic_entropynet_set_weights")
        self.weights = w
        return self

    def init_random(self, lh: int, lw: int, seed: int = 21, gain: float = 1.0, bias_range: float = 0.0):
        """Keras defaults (glorot-uniform kernels, zero biases) for a latent of lh x lw (fixes dense1's input size)."""
        self.set_weights(entropynet_glorot(lh, lw, seed, gain, bias_range))
        return self

    def __call__(self, x):
        lib, h = self.handle.lib, self.handle.h
        x, on_device = _f32_in(self.handle, x, "x")
        if x.ndim != 4 or x.shape[3] != 32:
            raise ValueError("expected float32 [P,h,w,32]")
        p, lh, lw, _ = x.shape
        if on_device:
            import torch
            out = torch.empty((p, 1), dtype=torch.float32, device=x.device)
            self.handle.check(lib.nnic_entropynet_forward(h, _ptr(x), p, lh, lw, _ptr(out), MEM_DEVICE, _stream_of(x)),
                              "nnic_entropynet_forward")
            return out
        out = np.empty((p, 1), np.float32)
        self.handle.check(lib.nnic_entropynet_forward(h, _ptr(x), p, lh, lw, _ptr(out), MEM_HOST, None), "nnic_entropynet_forward")
        return out


def entropynet_glorot(lh: int, lw: int, seed: int = 21, gain: float = 1.0, bias_range: float = 0.0) -> dict:
    """Glorot-uniform Entropynet weights for an lh x lw latent (Keras: limit sqrt(6 / (fan_in + fan_out)), Conv2D fans are
    k*k*Cin and k*k*Cout, Dense fans are its two sizes)."""
    rng = np.random.default_rng(seed)
    w = {}

    def draw(name, shape, fan_in, fan_out, nb):
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        w[name + "/kernel"] = (rng.uniform(-limit, limit, size=shape) * gain).astype(np.float32)
        w[name + "/bias"] = (rng.uniform(-bias_range, bias_range, size=(nb,)).astype(np.float32) if bias_range > 0
                             else np.zeros((nb,), np.float32))
    for name, shape in ENTROPYNET_CONVS:
        kh, kw, cin, cout = shape
        draw(name, shape, kh * kw * cin, kh * kw * cout, cout)
    feats = 64 * (-(-lh // 2)) * (-(-lw // 2))
    draw("dense1", (feats, 512), feats, 512, 512)
    draw("dense2", (512, 1), 512, 1, 1)
    return w


def noisy_quantise(handle: Handle, encoded, seed: int = 0, noise=None):
    """training.py:87-88: clip(encoded + U(-0.5, 0.5) / 255, 0, 1).  `noise` (same shape, values in [-0.5, 0.5)) replaces
    the generator (Philox4x32-10 keyed by `seed` and the element index)."""
    lib, h = handle.lib, handle.h
    x, on_device = _f32_in(handle, encoded, "encoded")
    n = None
    if noise is not None:
        n, n_dev = _f32_in(handle, noise, "noise")
        if n_dev != on_device or tuple(n.shape) != tuple(x.shape):
            raise ValueError("noise must have the shape of `encoded` and live on the same side of the bus")
    count = int(np.prod(x.shape))
    if on_device:
        import torch
        out = torch.empty_like(x)
        handle.check(lib.nnic_noise_quantise(h, _ptr(x), count, int(seed), _ptr(n), _ptr(out), MEM_DEVICE, _stream_of(x)),
                     "nnic_noise_quantise")
        return out
    out = np.empty_like(x)
    handle.check(lib.nnic_noise_quantise(h, _ptr(x), count, int(seed), _ptr(n), _ptr(out), MEM_HOST, None), "nnic_noise_quantise")
    return out


def ssim(handle: Handle, a, b):
    """tf.image.ssim(a, b, max_val=1.0) for single-channel images: a, b float32 [P,H,W,1] (or [P,H,W]) -> float32 [P]."""
    lib, h = handle.lib, handle.h
    a, a_dev = _f32_in(handle, a, "a")
    b, b_dev = _f32_in(handle, b, "b")
    if a_dev != b_dev or tuple(a.shape) != tuple(b.shape):
        raise ValueError("a and b must have the same shape and live on the same side of the bus")
    if a.ndim == 4:
        if a.shape[3] != 1:
            raise ValueError("expected single-channel images [P,H,W,1]")
    elif a.ndim != 3:
        raise ValueError("expected [P,H,W,1] or [P,H,W]")
    p, hh, ww = a.shape[:3]
    if hh < 11 or ww < 11:
        raise ValueError("tf.image.ssim needs images of at least 11 x 11")
    if a_dev:
        import torch
        out = torch.empty((p,), dtype=torch.float32, device=a.device)
        handle.check(lib.nnic_ssim(h, _ptr(a), _ptr(b), p, hh, ww, _ptr(out), MEM_DEVICE, _stream_of(a)), "nnic_ssim")
        return out
    out = np.empty((p,), np.float32)
    handle.check(lib.nnic_ssim(h, _ptr(a), _ptr(b), p, hh, ww, _ptr(out), MEM_HOST, None), "nnic_ssim")
    return out
