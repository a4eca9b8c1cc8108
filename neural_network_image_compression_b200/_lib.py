"""ctypes binding of libnnic.so (include/nnic.h).  There is no fallback: if the shared library is
missing or no sm_100 GPU is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NNIC_LIB: development override (e.g. the -DNNIC_TC_TIMERS build used by tools/); the product library sits next to this file
LIB_PATH = os.environ.get("NNIC_LIB") or os.path.join(_HERE, "libnnic.so")

MEM_HOST, MEM_DEVICE = 0, 1
ARITH_TC_SPLIT, ARITH_SIMT_F32 = 0, 1
ARITH_NAMES = {"tc_split": ARITH_TC_SPLIT, "simt_f32": ARITH_SIMT_F32}

# every symbol include/nnic.h declares: (name, restype, argtypes)
_u8p, _f32p, _u32p, _u64p, _vp = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint64), C.c_void_p)
SYMBOLS = (
    ("nnic_create", C.c_int, (C.c_int, C.POINTER(_vp))),
    ("nnic_destroy", None, (_vp,)),
    ("nnic_last_error", C.c_char_p, (_vp,)),
    ("nnic_version", C.c_char_p, ()),
    ("nnic_set_arith", C.c_int, (_vp, C.c_int)),
    ("nnic_get_arith", C.c_int, (_vp,)),
    ("nnic_launch_count", C.c_uint64, (_vp,)),
    ("nnic_set_weights", C.c_int, (_vp, C.c_int, C.c_int, _vp, _vp)),
    ("nnic_init_random", C.c_int, (_vp, C.c_int, C.c_uint64)),
    ("nnic_init_random_scaled", C.c_int, (_vp, C.c_int, C.c_uint64, C.c_double, C.c_double)),
    ("nnic_glorot_uniform", C.c_int, (C.c_int, C.c_uint64, C.c_double, C.c_double, _vp, _vp)),
    ("nnic_layer_shape", C.c_int, (C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))),
    ("nnic_encode", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp)),
    ("nnic_decode", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp)),
    ("nnic_run_encoder_planes", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_run_decoder_planes", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_rate", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp)),
    ("nnic_encode_rate", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp)),
    ("nnic_rate_channels", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_tensor_map_encodes", C.c_uint64, (_vp,)),
    ("nnic_hist_allreduce", C.c_int, (_vp, _vp, _vp, _vp)),
    ("nnic_set_decode_precision", C.c_int, (_vp, C.c_int)),
    ("nnic_get_decode_precision", C.c_int, (_vp,)),
    ("nnic_entropy_from_counts", C.c_int, (_vp, _vp, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_entropynet_set_weights", C.c_int, (_vp, C.c_int, _vp, _vp, C.c_int)),
    ("nnic_entropynet_forward", C.c_int, (_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_noise_quantise", C.c_int, (_vp, _vp, C.c_size_t, C.c_uint64, _vp, _vp, C.c_int, _vp)),
    ("nnic_ssim", C.c_int, (_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp)),
    ("nnic_set_micro_batch", C.c_int, (_vp, C.c_int)),
    ("nnic_scratch_bytes", C.c_size_t, (_vp,)),
    ("nnic_set_profiling", C.c_int, (_vp, C.c_int)),
    ("nnic_profile_collect", C.c_int, (_vp, _vp, _vp, C.c_int)),
    ("nnic_colour_constants", None, (_vp, _vp, _vp)),
    ("nnic_debug_fetch", C.c_longlong, (_vp, C.c_int, _vp, C.c_longlong)),
    ("nnic_debug_saturated", C.c_longlong, (_vp,)),
)

# include/nnic.h enum nnic_kernel_id
KERNEL_NAMES = ("conv1", "conv2", "conv3", "conv4", "conv8", "quantise", "latent_expand", "dconv1", "dconv5",
                "dconv6", "dconv7", "dconv8", "hist", "entropy", "hist_reduce", "f32_split", "entropynet_conv", "dense", "ssim", "noise")

_lib = None


class NnicError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libnnic.so and bind every declared symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NnicError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = list(argtypes)
    _lib = lib
    return lib


def _ptr(a):
    """Address of a NumPy array / torch tensor / int (device pointer) / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


class Handle:
    """One native codec instance bound to one GPU (nnic_t)."""

    def __init__(self, device: int = 0, arith: str = "tc_split"):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.nnic_create(int(device), C.byref(h))
        if rc != 0:
            raise NnicError(f"nnic_create(device={device}) failed ({rc}): {self.lib.nnic_last_error(None).decode()}")
        self.h = h
        self.device = int(device)
        self.set_arith(arith)

    def close(self):
        if getattr(self, "h", None):
            self.lib.nnic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            raise NnicError(f"{what} failed ({rc}): {self.lib.nnic_last_error(self.h).decode()}")

    def set_arith(self, arith):
        code = ARITH_NAMES[arith] if isinstance(arith, str) else int(arith)
        self.check(self.lib.nnic_set_arith(self.h, code), "nnic_set_arith")

    def set_decode_precision(self, precision):
        """'split' (default, as exact as the encoder) or 'fp16' (one fp16 product per MAC in the decoder; see
        include/nnic.h enum nnic_decode_precision)."""
        code = {"split": 0, "fp16": 1}[precision] if isinstance(precision, str) else int(precision)
        self.check(self.lib.nnic_set_decode_precision(self.h, code), "nnic_set_decode_precision")

    @property
    def decode_precision(self) -> str:
        return ("split", "fp16")[self.lib.nnic_get_decode_precision(self.h)]

    @property
    def arith(self) -> str:
        code = self.lib.nnic_get_arith(self.h)
        return {v: k for k, v in ARITH_NAMES.items()}[code]

    @property
    def launch_count(self) -> int:
        return int(self.lib.nnic_launch_count(self.h))

    def set_micro_batch(self, n: int):
        self.check(self.lib.nnic_set_micro_batch(self.h, int(n)), "nnic_set_micro_batch")

    def set_weights(self, set_index: int, layer: int, kernel: np.ndarray, bias: np.ndarray):
        k = np.ascontiguousarray(kernel, np.float32)
        b = np.ascontiguousarray(bias, np.float32)
        self.check(self.lib.nnic_set_weights(self.h, set_index, layer, _ptr(k), _ptr(b)), "nnic_set_weights")

    def set_profiling(self, on: bool):
        self.check(self.lib.nnic_set_profiling(self.h, int(bool(on))), "nnic_set_profiling")

    def profile_collect(self) -> dict:
        """{kernel name: (summed ms, launches)} since the previous collect (profiling must be on)."""
        ms = np.zeros(len(KERNEL_NAMES), np.float32)
        cnt = np.zeros(len(KERNEL_NAMES), np.int32)
        rc = self.lib.nnic_profile_collect(self.h, _ptr(ms), _ptr(cnt), len(KERNEL_NAMES))
        if rc < 0:
            self.check(rc, "nnic_profile_collect")
        return {KERNEL_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(len(KERNEL_NAMES)) if cnt[i]}

    def saturated_activations(self) -> int:
        """Values of the last encode / decode micro-batch that hit the fp16 limit of the split representation (|v| > 4094)."""
        n = int(self.lib.nnic_debug_saturated(self.h))
        if n < 0:
            self.check(n, "nnic_debug_saturated")
        return n

    def debug_fetch(self, slot: int) -> np.ndarray:
        n = self.lib.nnic_debug_fetch(self.h, slot, None, 0)
        if n < 0:
            self.check(int(n), "nnic_debug_fetch")
        out = np.empty(int(n), np.float32)
        if n:
            got = self.lib.nnic_debug_fetch(self.h, slot, _ptr(out), n)
            if got < 0:
                self.check(int(got), "nnic_debug_fetch")
        return out


def colour_constants():
    lib = load_library()
    k, kinv, off = np.empty((3, 3), np.float32), np.empty((3, 3), np.float32), np.empty(3, np.float32)
    lib.nnic_colour_constants(_ptr(k), _ptr(kinv), _ptr(off))
    return k, kinv, off
