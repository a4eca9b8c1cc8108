"""Generate the full-size parity records tests/golden/parity_<config>_<weights>.npz (see tests/parity_fixture.py) and the
verbatim copy of the reference's ImageNet patches 00000..04095 that config 3 is defined on.

Run in the build container (needs /root/reference for the patches; everything else is rebuilt from committed data):
    python tools/make_parity_fixtures.py [c2 c3 c4 c5] [--weights default spread]
The oracle (oracle/nnic_oracle.py, fp64 and fp32 modes) is run at the FULL size of every BASELINE.json configuration:
    c2  24 x 512x768     encode + rate + decode
    c3  4096 x 128x128   encode + rate                 (the reference's real patches)
    c4  16 x 2160x3840   encode, then decode of the fp64-oracle latent
    c5  patches [0, 2048) of the 65 536-patch set of bench.py's batch-sharded workload, encode + rate
This pins the CUDA path to the ORACLE at full size; the oracle itself stays "parity unpinned" against TensorFlow.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_fixture as PF
from conftest import make_weights
from oracle import nnic_oracle as O

REF_PATCHES = "/root/reference/data/imagenet_patches"


def copy_c3_patches(count=4096):
    path = os.path.join(PF.GOLDEN, "imagenet_patches_c3.npz")
    if os.path.exists(path):
        return
    blobs, offs = [], [0]
    for i in range(count):
        with open(os.path.join(REF_PATCHES, f"{i:05d}.jpg"), "rb") as f:
            b = f.read()
        blobs.append(np.frombuffer(b, np.uint8))
        offs.append(offs[-1] + len(b))
    np.savez(path, jpeg_bytes=np.concatenate(blobs), offsets=np.array(offs, np.int64))
    print("wrote", path, offs[-1], "bytes of JPEG data")


def make(config, wname, images, chunk, with_decode, H, W):
    t0 = time.time()
    eY, eC, dY, dC = make_weights(wname)
    lat_b, rec_b = PF.RecordBuilder(), PF.RecordBuilder()
    n = images.shape[0]
    hist = np.zeros((3, 256), np.int64)
    bpp64, psnr64 = [], []
    for i0 in range(0, n, chunk):
        img = images[i0:i0 + chunk]
        pre64 = O.encode_prequant(img, eY, eC, "f64")
        sym32 = O.encode(img, eY, eC, "f32")
        sym64 = lat_b.add(pre64 * 255.0, sym32)
        h_, _ent, bpp, hg = O.rate(sym64, H, W, "f32")
        hist += hg
        bpp64.append(bpp)
        if with_decode:
            d64 = O.decode_prequant(sym64, dY, dC, "f64")
            rec32 = O.decode(sym64, dY, dC, "f32")
            rec64 = rec_b.add(d64 * 255.0, rec32)
            psnr64.extend(O.psnr(img[k], rec64[k]) for k in range(img.shape[0]))
        print(f"  {config} {wname}: {min(i0 + chunk, n)}/{n} images, {time.time() - t0:.0f} s", flush=True)
    out = {"input_sha1": np.array(PF.sha1(images)), "shape": np.array(images.shape, np.int64)}
    out.update(lat_b.finish("lat"))
    out["lat_hist64"] = hist
    out["lat_bpp64"] = np.concatenate(bpp64).astype(np.float32)
    if with_decode:
        out.update(rec_b.finish("rec"))
        out["rec_psnr64"] = np.array(psnr64, np.float64)
    path = os.path.join(PF.GOLDEN, f"parity_{config}_{wname}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; lat ties {out['lat_tie_delta'].size} of {n * lat_b.per_image} "
          f"(f32 oracle vs f64 oracle: {lat_b.mism32} mismatches, {lat_b.mism32_outside} outside the band)"
          + (f"; rec ties {out['rec_tie_delta'].size} (f32 vs f64: {rec_b.mism32}, outside {rec_b.mism32_outside})" if with_decode else "")
          + f"; {time.time() - t0:.0f} s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c2", "c3", "c5", "c4"])
    ap.add_argument("--weights", nargs="*", default=["spread", "default"])
    args = ap.parse_args()
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    for cfg in args.configs:
        if cfg == "c2":
            images, chunk, dec, H, W = PF.c2_images(), 4, True, 512, 768
        elif cfg == "c3":
            copy_c3_patches()
            images, chunk, dec, H, W = PF.c3_patches(), 256, False, 128, 128
        elif cfg == "c4":
            images, chunk, dec, H, W = PF.c4_images(), 1, True, 2160, 3840
        elif cfg == "c5":
            images, chunk, dec, H, W = PF.c5_patches(0, 2048).numpy(), 128, False, 256, 256
        else:
            raise SystemExit(f"unknown config {cfg}")
        for wname in args.weights:
            make(cfg, wname, images, chunk, dec, H, W)


if __name__ == "__main__":
    main()
