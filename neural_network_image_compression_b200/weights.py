"""Weight containers for the four networks of the codec (encoder Y / CbCr, decoder Y / CbCr).

Layer tables follow /root/reference/tf2_0/src/encoder.py:10-17 and decoder.py:10-17; array layouts
are the Keras ones the reference's checkpoints hold (SURVEY.md 8b "Weights"):
  Conv2D           kernel [kh, kw, Cin, Cout], bias [Cout]
  Conv2DTranspose  kernel [kh, kw, Cout, Cin], bias [Cout]
A weight set is a dict {'<layer>/kernel': float32 array, '<layer>/bias': float32 array}.
"""
from __future__ import annotations

import numpy as np

# (name, ksize, stride, cin, cout)
ENCODER_LAYERS = (("conv1", 5, 2, 1, 32), ("conv2", 5, 2, 32, 64), ("conv3", 3, 1, 64, 64),
                  ("conv4", 3, 1, 64, 64), ("conv8", 5, 2, 64, 32))
DECODER_LAYERS = (("dconv1", 5, 2, 32, 64), ("dconv5", 3, 1, 64, 64), ("dconv6", 3, 1, 64, 64),
                  ("dconv7", 5, 2, 64, 64), ("dconv8", 5, 2, 64, 1))

# index of each network inside the native handle (include/nnic.h NNIC_SET_*)
SET_ENC_Y, SET_ENC_CBCR, SET_DEC_Y, SET_DEC_CBCR = 0, 1, 2, 3
# the reference's checkpoint suffixes (tf2_0/src/utils.py:5, 26-28)
MODEL_SUFFIXES = ("Y", "CbCr")

DEFAULT_SEEDS = {"encY": 11, "encCbCr": 12, "decY": 13, "decCbCr": 14}


def kernel_shape(kind: str, k: int, cin: int, cout: int):
    return (k, k, cin, cout) if kind == "encoder" else (k, k, cout, cin)


def layers_of(kind: str):
    return ENCODER_LAYERS if kind == "encoder" else DECODER_LAYERS


def glorot_uniform(kind: str, seed: int, gain: float = 1.0, bias_range: float = 0.0) -> dict:
    """Keras default initialisation (glorot-uniform kernel, zero bias) for one network.

    limit = sqrt(6 / ((Cin + Cout) * kh * kw)) for both kernel layouts.  `gain` scales the kernels and
    `bias_range` draws biases from U(-r, r); (1, 0) is the Keras default ("W-default" in SURVEY 8d),
    other values give the "W-spread" sets that exercise both clamps and every histogram bin.
    """
    rng = np.random.default_rng(seed)
    out = {}
    for name, k, _s, cin, cout in layers_of(kind):
        limit = np.sqrt(6.0 / ((cin + cout) * k * k))
        shape = kernel_shape(kind, k, cin, cout)
        out[name + "/kernel"] = (rng.uniform(-limit, limit, size=shape) * gain).astype(np.float32)
        if bias_range > 0:
            out[name + "/bias"] = rng.uniform(-bias_range, bias_range, size=(cout,)).astype(np.float32)
        else:
            out[name + "/bias"] = np.zeros((cout,), np.float32)
    return out


def glorot_uniform_native(kind: str, seed: int, gain: float = 1.0, bias_range: float = 0.0, set_index: int | None = None) -> dict:
    """The same network drawn by libnnic.so (nnic_glorot_uniform: PCG64 / SeedSequence restated in C++; host code, no
    GPU needed).  Bit-identical to glorot_uniform() -- tests/test_host.py::test_c_glorot_matches_numpy."""
    from ._lib import load_library
    lib = load_library()
    if set_index is None:
        set_index = SET_ENC_Y if kind == "encoder" else SET_DEC_Y
    layers = layers_of(kind)
    nk = sum(k * k * cin * cout for _n, k, _s, cin, cout in layers)
    nb = sum(cout for *_r, cout in layers)
    kern, bias = np.empty(nk, np.float32), np.empty(nb, np.float32)
    rc = lib.nnic_glorot_uniform(int(set_index), int(seed), float(gain), float(bias_range), kern.ctypes.data, bias.ctypes.data)
    if rc != 0:
        raise ValueError(f"nnic_glorot_uniform failed ({rc})")
    out, ko, bo = {}, 0, 0
    for name, k, _s, cin, cout in layers:
        n = k * k * cin * cout
        out[name + "/kernel"] = kern[ko:ko + n].reshape(kernel_shape(kind, k, cin, cout)).copy()
        out[name + "/bias"] = bias[bo:bo + cout].copy()
        ko += n
        bo += cout
    return out


def check_weight_set(kind: str, w: dict) -> None:
    for name, k, _s, cin, cout in layers_of(kind):
        kern, bias = w[name + "/kernel"], w[name + "/bias"]
        if tuple(kern.shape) != kernel_shape(kind, k, cin, cout):
            raise ValueError(f"{name}/kernel has shape {kern.shape}, expected {kernel_shape(kind, k, cin, cout)}")
        if tuple(bias.shape) != (cout,):
            raise ValueError(f"{name}/bias has shape {bias.shape}, expected {(cout,)}")


def save_npz(path: str, w: dict) -> None:
    np.savez(path, **{k.replace("/", "__"): v for k, v in w.items()})


def load_npz(path: str) -> dict:
    with np.load(path) as z:
        return {k.replace("__", "/"): np.asarray(z[k], np.float32) for k in z.files}
