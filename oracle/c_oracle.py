"""ctypes front end of oracle/nnic_oracle.c (plain-C restatement, fp64).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnnic_oracle_c.so")
ENC_LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv8")
DEC_LAYERS = ("dconv1", "dconv5", "dconv6", "dconv7", "dconv8")


def load():
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(LIB_PATH)
    pp = C.POINTER(C.c_void_p)
    lib.oracle_c_encode_prequant.argtypes = [C.c_void_p, C.c_int, C.c_int, pp, pp, C.c_void_p, C.c_int, C.c_int]
    lib.oracle_c_decode_prequant.argtypes = [C.c_void_p, C.c_int, C.c_int, pp, pp, C.c_void_p]
    lib.oracle_c_quantise.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    return lib


def _ptrs(w: dict, layers):
    keep = []
    arr = (C.c_void_p * 10)()
    for i, name in enumerate(layers):
        for j, var in enumerate(("kernel", "bias")):
            a = np.ascontiguousarray(w[f"{name}/{var}"], np.float32)
            keep.append(a)
            arr[2 * i + j] = a.ctypes.data
    return arr, keep


def encode_prequant(img_u8: np.ndarray, enc_y: dict, enc_cbcr: dict) -> np.ndarray:
    """uint8 [N,H,W,3] -> float64 [N,ceil(H/8),ceil(W/8),96] (the clipped encoder output before *255/round)."""
    lib = load()
    n, hh, ww, _ = img_u8.shape
    h, w = -(-(-(-(-(-hh // 2)) // 2)) // 2), -(-(-(-(-(-ww // 2)) // 2)) // 2)
    py, ky = _ptrs(enc_y, ENC_LAYERS)
    pc, kc = _ptrs(enc_cbcr, ENC_LAYERS)
    out = np.empty((n, h, w, 96), np.float64)
    for i in range(n):
        x = np.ascontiguousarray(img_u8[i])
        rc = lib.oracle_c_encode_prequant(x.ctypes.data, hh, ww, py, pc, out[i].ctypes.data, h, w)
        assert rc == 0
    return out


def decode_prequant(latent_u8: np.ndarray, dec_y: dict, dec_cbcr: dict) -> np.ndarray:
    """uint8 [N,h,w,96] -> float64 [N,8h,8w,3] (clipped RGB in [0,1] before *255/round)."""
    lib = load()
    n, h, w, _ = latent_u8.shape
    py, ky = _ptrs(dec_y, DEC_LAYERS)
    pc, kc = _ptrs(dec_cbcr, DEC_LAYERS)
    out = np.empty((n, 8 * h, 8 * w, 3), np.float64)
    for i in range(n):
        x = np.ascontiguousarray(latent_u8[i])
        assert lib.oracle_c_decode_prequant(x.ctypes.data, h, w, py, pc, out[i].ctypes.data) == 0
    return out


def quantise(v: np.ndarray) -> np.ndarray:
    lib = load()
    v = np.ascontiguousarray(v, np.float64)
    out = np.empty(v.shape, np.uint8)
    lib.oracle_c_quantise(v.ctypes.data, v.size, out.ctypes.data)
    return out
