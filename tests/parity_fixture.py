"""Compact full-size parity records of the oracle (TEST INFRASTRUCTURE).

The oracle needs minutes of CPU time at the sizes of BASELINE.json configs 3-5 (and the fp64 mode several times that),
which the GPU tests cannot afford on every run.  `tools/make_parity_fixtures.py` therefore runs the oracle ONCE, in the
build container, at the full size of every configuration and stores a record from which the comparison "CUDA output vs
oracle output" can be repeated exactly without the oracle's output tensor:

  * digest[n]     blake2b-64 of image n's fp64-oracle bytes with every position inside the rounding-tie band zeroed;
                  equal digests <=> every byte OUTSIDE the band is bit-identical to the fp64 oracle;
  * tie_idx       flat positions whose pre-rounding fp64 value lies within TIE_BAND of k + 1/2 (the documented ties:
                  the only places where two fp32-accurate arithmetics may legitimately round differently);
  * tie_sym64 / tie_sym32 / tie_floor   the fp64-oracle byte, the fp32-oracle byte and floor(value) at those positions
                  (a byte at a tie must be floor or floor + 1);
  * hist64        the global [3,256] symbol histogram of the fp64-oracle latent, bpp64[n], psnr64[n] where they apply.

Live-oracle tests (tests/test_oracle.py, CPU) recompute random images of every record and compare, so the records
cannot drift from oracle/nnic_oracle.py unnoticed.  Inputs are rebuilt from committed data by the functions below
(the same functions the generator used); their SHA-1 is part of the record.
"""
from __future__ import annotations

import hashlib
import io
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TIE_BAND = 2e-3          # symbol units, BASELINE.json "documented rounding ties" 
MISMATCH_LIMIT = 1e-4


# ---------------------------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------------------------
def kodim21() -> np.ndarray:
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, "kodim21.png")))


def c2_images() -> np.ndarray:
    """image 0 = kodim21; images 1-11 = flips and 8-pixel-multiple cyclic shifts of it; images 12-23 =
    mosaics (4 x 6 tiles of 128 x 128) of the golden ImageNet patches with per-tile flips.  Fixed recipe, seed 0."""
    k = kodim21()
    with np.load(os.path.join(GOLDEN, "imagenet_patches.npz")) as z:
        patches = z["input"]
    rng = np.random.default_rng(0)
    imgs = [k, k[::-1], k[:, ::-1], k[::-1, ::-1]]
    while len(imgs) < 12:
        dy, dx = 8 * int(rng.integers(1, 64)), 8 * int(rng.integers(1, 96))
        imgs.append(np.roll(imgs[len(imgs) % 4], (dy, dx), axis=(0, 1)))
    while len(imgs) < 24:
        m = np.empty((512, 768, 3), np.uint8)
        for ty in range(4):
            for tx in range(6):
                t = patches[int(rng.integers(0, patches.shape[0]))]
                if rng.integers(0, 2):
                    t = t[::-1]
                if rng.integers(0, 2):
                    t = t[:, ::-1]
                m[128 * ty:128 * ty + 128, 128 * tx:128 * tx + 128] = t
        imgs.append(m)
    return np.ascontiguousarray(np.stack(imgs))


def c3_patches(count: int = 4096) -> np.ndarray:
    """the reference's own patches data/imagenet_patches/00000..04095.jpg (128 x 128), stored verbatim
    (JPEG bytes) in tests/golden/imagenet_patches_c3.npz by tools/make_parity_fixtures.py."""
    from PIL import Image
    with np.load(os.path.join(GOLDEN, "imagenet_patches_c3.npz")) as z:
        blob, offs = z["jpeg_bytes"], z["offsets"]
    out = np.empty((count, 128, 128, 3), np.uint8)
    raw = blob.tobytes()
    for i in range(count):
        out[i] = np.array(Image.open(io.BytesIO(raw[offs[i]:offs[i + 1]])).convert("RGB"))
    return out


def c4_images(count: int = 16) -> np.ndarray:
    """3840 x 2160 images built by tiling kodim21 (5 x 4.2 tiles, cropped) with per-image flips and
    an 8-pixel-multiple cyclic shift, so that the 16 images differ."""
    k = kodim21()
    tile = np.tile(k, (5, 5, 1))[:2160, :3840]
    out = np.empty((count, 2160, 3840, 3), np.uint8)
    for i in range(count):
        t = tile
        if i & 1:
            t = t[::-1]
        if i & 2:
            t = t[:, ::-1]
        out[i] = np.roll(t, (8 * 17 * (i // 4), 8 * 29 * (i // 4)), axis=(0, 1))
    return out


def c5_patches(first: int, count: int, device="cpu", h: int = 256, w: int = 256, salt: int = 0):
    """256 x 256 patches whose bytes depend only on (global patch index, position, salt) -- integer
    hash, identical on CPU and GPU and for every sharding of the set.  Same generator as bench.py (sharded_patches_gpu)."""
    import sys
    import torch
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    return bench.sharded_patches_gpu(torch, first, count, h, w, salt, torch.device(device))


def sha1(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------------------------------------------
# records
# ---------------------------------------------------------------------------------------------------------------
def _digest(img_bytes: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(img_bytes).tobytes(), digest_size=8).digest(), np.uint64)[0]


class RecordBuilder:
    """Accumulates the record of one output tensor (latent or reconstruction) image by image."""

    def __init__(self):
        self.per_image = 0
        self.n = 0
        self.digests, self.idx, self.s64, self.s32, self.flo = [], [], [], [], []
        self.mism32 = 0
        self.mism32_outside = 0

    def add(self, scaled64: np.ndarray, sym32: np.ndarray):
        """scaled64: fp64 pre-rounding values * 255, [n, ...]; sym32: the fp32 oracle's bytes, same shape."""
        n = scaled64.shape[0]
        per = int(np.prod(scaled64.shape[1:]))
        assert self.per_image in (0, per)
        self.per_image = per
        s = scaled64.reshape(n, per)
        sym64 = np.round(s).astype(np.uint8)
        s32 = sym32.reshape(n, per)
        fl = np.floor(s)
        tie = np.abs(s - fl - 0.5) < TIE_BAND
        diff32 = sym64 != s32
        self.mism32 += int(diff32.sum())
        self.mism32_outside += int((diff32 & ~tie).sum())
        for i in range(n):
            pos = np.flatnonzero(tie[i])
            masked = sym64[i].copy()
            masked[pos] = 0
            self.digests.append(_digest(masked))
            self.idx.append(pos.astype(np.int64) + (self.n + i) * per)
            self.s64.append(sym64[i][pos]); self.s32.append(s32[i][pos]); self.flo.append(fl[i][pos].astype(np.uint8))
        self.n += n
        return sym64.reshape(scaled64.shape)

    def finish(self, prefix: str) -> dict:
        idx = np.concatenate(self.idx) if self.idx else np.zeros(0, np.int64)
        delta = np.diff(idx, prepend=0)
        assert delta.size == 0 or delta.max() < 2 ** 32
        return {f"{prefix}_digest": np.array(self.digests, np.uint64), f"{prefix}_tie_delta": delta.astype(np.uint32),
                f"{prefix}_tie_sym64": np.concatenate(self.s64) if self.s64 else np.zeros(0, np.uint8),
                f"{prefix}_tie_sym32": np.concatenate(self.s32) if self.s32 else np.zeros(0, np.uint8),
                f"{prefix}_tie_floor": np.concatenate(self.flo) if self.flo else np.zeros(0, np.uint8),
                f"{prefix}_meta": np.array([self.n, self.per_image, self.mism32, self.mism32_outside], np.int64)}


def load_record(name: str) -> dict:
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def tie_positions(rec: dict, prefix: str) -> np.ndarray:
    return np.cumsum(rec[f"{prefix}_tie_delta"].astype(np.int64))


def compare(got: np.ndarray, rec: dict, prefix: str, first_image: int = 0) -> dict:
    """Compare CUDA bytes `got` [n, ...] (images first_image .. first_image + n of the record) with the record.
    Returns the statistics the parity report carries; raises AssertionError when a byte outside the tie band differs,
    when a tie byte is neither floor nor floor + 1, or when the mismatch fraction exceeds MISMATCH_LIMIT."""
    n_total, per, mism32, mism32_out = (int(v) for v in rec[f"{prefix}_meta"])
    n = got.shape[0]
    assert got.dtype == np.uint8 and int(np.prod(got.shape[1:])) == per and first_image + n <= n_total
    flat = got.reshape(n, per)
    pos_all = tie_positions(rec, prefix)
    lo, hi = np.searchsorted(pos_all, [first_image * per, (first_image + n) * per])
    pos = pos_all[lo:hi] - first_image * per
    s64, s32, fl = (rec[f"{prefix}_tie_{k}"][lo:hi] for k in ("sym64", "sym32", "floor"))
    at = flat.reshape(-1)[pos]
    masked = flat.copy()
    masked.reshape(-1)[pos] = 0
    bad_images = [first_image + i for i in range(n) if _digest(masked[i]) != rec[f"{prefix}_digest"][first_image + i]]
    legal = (at == fl) | (at.astype(np.int16) == fl.astype(np.int16) + 1)
    stats = {"values": int(flat.size), "tie_band_positions": int(pos.size),
             "mismatch_vs_f64": int((at != s64).sum()), "mismatch_vs_f32": int((at != s32).sum()),
             "mismatch_outside_tie_band": 0 if not bad_images else None,
             "f32_oracle_vs_f64_oracle": mism32 if n == n_total else None,
             "mismatch_fraction_vs_f64": float((at != s64).sum()) / flat.size,
             "_flips_per_image": np.bincount(pos[at != s64] // per, minlength=n)}
    assert not bad_images, f"{prefix}: bytes outside the rounding-tie band differ from the fp64 oracle in images {bad_images[:8]}"
    assert bool(legal.all()), f"{prefix}: {int((~legal).sum())} tie bytes are neither floor nor floor+1"
    assert stats["mismatch_fraction_vs_f64"] <= MISMATCH_LIMIT, f"{prefix}: mismatch fraction {stats['mismatch_fraction_vs_f64']:.2e}"
    assert mism32_out == 0
    return stats


def oracle_bytes(got: np.ndarray, rec: dict, prefix: str, first_image: int = 0) -> np.ndarray:
    """`got` with its tie positions replaced by the fp64 oracle's bytes: after compare() passed this IS the oracle's
    output tensor (used to feed the decoder exactly the latent the decode record was generated from)."""
    _n_total, per, _a, _b = (int(v) for v in rec[f"{prefix}_meta"])
    n = got.shape[0]
    pos_all = tie_positions(rec, prefix)
    lo, hi = np.searchsorted(pos_all, [first_image * per, (first_image + n) * per])
    out = got.copy()
    out.reshape(-1)[pos_all[lo:hi] - first_image * per] = rec[f"{prefix}_tie_sym64"][lo:hi]
    return out


def histogram_from_record(got_latent: np.ndarray, rec: dict, prefix: str) -> np.ndarray:
    """The [3,256] histogram the oracle's latent has, adjusted by the documented tie flips of `got_latent`:
    what an exact histogram of `got_latent` must equal."""
    hist = rec[f"{prefix}_hist64"].astype(np.int64).copy()
    _n_total, per, _a, _b = (int(v) for v in rec[f"{prefix}_meta"])
    pos = tie_positions(rec, prefix)
    at = got_latent.reshape(-1)[pos]
    s64 = rec[f"{prefix}_tie_sym64"]
    flip = at != s64
    plane = (pos[flip] % 96) // 32
    np.subtract.at(hist, (plane, s64[flip]), 1)
    np.add.at(hist, (plane, at[flip]), 1)
    return hist


def verify_images_with_live_oracle(rec: dict, prefix: str, image_indices, scaled64: np.ndarray, sym32: np.ndarray | None = None):
    """CPU check that the record describes what the oracle produces NOW: `scaled64` are the live fp64 oracle's
    pre-rounding values * 255 of the images `image_indices`; their digest, tie positions and tie bytes must be the
    record's (and the fp32 oracle's bytes at the ties, when given)."""
    n_total, per, _a, _b = (int(v) for v in rec[f"{prefix}_meta"])
    pos_all = tie_positions(rec, prefix)
    s = scaled64.reshape(len(image_indices), per)
    for k, i in enumerate(image_indices):
        sym64 = np.round(s[k]).astype(np.uint8)
        fl = np.floor(s[k])
        pos = np.flatnonzero(np.abs(s[k] - fl - 0.5) < TIE_BAND)
        lo, hi = np.searchsorted(pos_all, [i * per, (i + 1) * per])
        assert np.array_equal(pos_all[lo:hi] - i * per, pos), f"{prefix}: tie positions of image {i} changed"
        assert np.array_equal(rec[f"{prefix}_tie_sym64"][lo:hi], sym64[pos])
        assert np.array_equal(rec[f"{prefix}_tie_floor"][lo:hi], fl[pos].astype(np.uint8))
        if sym32 is not None:
            assert np.array_equal(rec[f"{prefix}_tie_sym32"][lo:hi], sym32.reshape(len(image_indices), per)[k][pos])
        masked = sym64.copy()
        masked[pos] = 0
        assert _digest(masked) == rec[f"{prefix}_digest"][i], f"{prefix}: digest of image {i} changed"
