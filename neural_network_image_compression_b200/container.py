"""File side of the codec: the callers of the hot path in /root/reference/tf2_0/src/utils.py:30-62,85-120.

The reference stores a compressed image as an RGB PNG whose three channels are the three 32-channel groups of the
latent, each viewed as a 4h x 8w byte image (a plain C-order reshape of [h,w,32], `_feed_batch`, utils.py:35-44);
`uncompress` undoes the reshape and runs the decoder.  Everything here is host plumbing around `Encoder()(x)` /
`Decoder()(x)`; the arithmetic stays in libnnic.so.
"""
from __future__ import annotations

import io
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from PIL import Image

IMAGE_EXTENSIONS = ("png", "jpg", "jpeg", "gif", "pgm", "ppm", "bmp", "jp2")   # utils.py:94


def pack_latent(latent: np.ndarray) -> np.ndarray:
    """uint8 [n,h,w,96] -> uint8 [n,4h,8w,3] (utils.py:42-44): channel i of the picture is the C-order byte
    stream of latent[..., 32i:32i+32] cut into rows of 8w bytes."""
    n, h, w, c = latent.shape
    if c != 96:
        raise ValueError("latent must have 96 channels")
    return np.stack([np.ascontiguousarray(latent[..., 32 * i:32 * (i + 1)]).reshape(n, 4 * h, 8 * w) for i in range(3)],
                    axis=3)


def unpack_latent(picture: np.ndarray) -> np.ndarray:
    """uint8 [n,4h,8w,3] -> uint8 [n,h,w,96] (utils.py:36-38), the inverse of pack_latent."""
    n, hh, ww, c = picture.shape
    if c != 3 or hh % 4 or ww % 8:
        raise ValueError("a packed latent is an RGB picture of 4h x 8w pixels")
    return np.concatenate([np.ascontiguousarray(picture[..., i]).reshape(n, hh // 4, ww // 8, 32) for i in range(3)],
                          axis=3)


def save_img(img: np.ndarray, output_dir: str, filename: str) -> str:
    """utils.py:85-87: integer-valued array -> `<output_dir>/<filename>.png`, written with optimize=True."""
    if not np.array_equal(np.round(img), img):
        raise AssertionError("save_img expects integer-valued pixels")
    path = os.path.join(output_dir, filename + ".png")
    Image.fromarray(np.asarray(img, np.uint8)).save(path, optimize=True)
    return path


def read_dataset(dataset_path: str):
    """utils.py:89-120: the colour images of a directory in sorted file-name order (grey-scale files are skipped,
    as in the reference) and their names without extension.

    Returns (images, filenames): images is a uint8 array [N,H,W,C] when all files have one size, else a list of
    [1,H,W,C] arrays (the reference builds an object array there and then fails to batch it; a list of
    single-image batches is what its batch_size = 1 branch intends)."""
    imgs, names = [], []
    for f in sorted(os.listdir(dataset_path)):
        if f.rsplit(".", 1)[-1] not in IMAGE_EXTENSIONS:
            continue
        with Image.open(os.path.join(dataset_path, f)) as im:
            arr = np.array(im)
        if arr.ndim == 3:
            imgs.append(arr.astype(np.uint8))
            names.append(f.rsplit(".", 1)[0])
    if imgs and all(a.shape == imgs[0].shape for a in imgs):
        return np.stack(imgs), names
    return [a[None] for a in imgs], names


def png_size(picture: np.ndarray, compress_level: int = 6) -> int:
    """Bytes of the PNG encoding of one uint8 picture ([H,W], [H,W,1] or [H,W,3])."""
    buf = io.BytesIO()
    Image.fromarray(np.squeeze(picture)).save(buf, format="PNG", compress_level=compress_level)
    return buf.getbuffer().nbytes


def get_bpp(encoded: np.ndarray, tot_pixels_compressed: float | None = None) -> np.ndarray:
    """tf2_0/src/training.py:14-21: `encoded` float [P,h,w,32] in 0..255 (one row per colour plane) -> float32 [P,1],
    8 * PNG bytes of the [4h,8w] byte picture / tot_pixels_compressed (default: the picture's own 4h*8w pixels).

    The reference measures the size with tf.image.encode_png (libpng, zlib default level); this uses Pillow at the
    same zlib level.  Byte counts of two PNG encoders differ by their filter heuristics, so this figure is NOT part
    of the parity contract."""
    enc = np.asarray(encoded)
    p, h, w, c = enc.shape
    if c != 32:
        raise ValueError("expected [P,h,w,32]")
    pics = np.round(enc).astype(np.uint8).reshape(p, 4 * h, 8 * w)
    tot = float(4 * h * 8 * w) if tot_pixels_compressed is None else float(tot_pixels_compressed)
    sizes = np.array([png_size(pic) for pic in pics], np.float32)
    return (8.0 * sizes / np.float32(tot)).reshape(-1, 1).astype(np.float32)


class DatasetDriver:
    """Mixin of ProClass: `_feed_batch` / `_use_model` of utils.py:30-62."""

    batch_size = 4        # utils.py:57

    def _feed_batch(self, x, filenames, output_dir, in_cshape, pool=None):
        x = np.asarray(x)
        if x.ndim == 5:
            x = x[0]
        if in_cshape == 96 and x.shape[3] == 3:
            x = unpack_latent(x)
        out = self(x)                                  # Encoder.__call__ / Decoder.__call__ on the GPU
        if out.shape[3] == 96:
            out = pack_latent(out)
        jobs = []
        for i in range(out.shape[0]):
            if pool is None:
                save_img(np.squeeze(out[i]), output_dir, filenames[i])
            else:
                jobs.append(pool.submit(save_img, np.squeeze(out[i]), output_dir, filenames[i]))
        return jobs

    def _use_model(self, dataset_path, checkpoint_path, output_dir, in_cshape, threads: int | None = None):
        """Load the weights (checkpoint_path=None keeps the installed ones), run every image of `dataset_path`
        through the model in batches and write one PNG per image to `output_dir`.  PNG encoding (zlib, the slow part)
        runs on a thread pool while the next batch is on the GPU."""
        os.makedirs(output_dir, exist_ok=True)
        if checkpoint_path is not None:
            self.load(checkpoint_path)
        x, filenames = read_dataset(dataset_path)
        tot_n = len(filenames)
        ragged = isinstance(x, list)
        bs = 1 if ragged else self.batch_size
        with ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 1)) as pool:
            jobs = []
            for lo in range(0, tot_n, bs):
                hi = min(tot_n, lo + bs)
                batch = x[lo] if ragged else x[lo:hi]
                jobs += self._feed_batch(batch, filenames[lo:hi], output_dir, in_cshape, pool)
            for j in jobs:
                j.result()
        return output_dir
