"""Generate tests/golden/*.npz (run in the build container, where /root/reference exists).

Inputs are real data shipped by the reference (data/kodak_img/kodim21.png, data/imagenet_patches/*.jpg);
outputs are produced by oracle/nnic_oracle.py in both modes.  The reference's own implementation cannot
run here (TensorFlow is not installed), so these vectors pin the ORACLE, not TensorFlow -- "parity unpinned"
in the sense of SURVEY.md 8c; they keep the oracle and the CUDA path from drifting apart silently.
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_network_image_compression_b200 import weights as Wt
from oracle import nnic_oracle as O

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")

WEIGHT_SETS = {"default": (1.0, 0.0), "spread": (1.6, 0.05)}


def weight_sets(name):
    gain, br = WEIGHT_SETS[name]
    return (Wt.glorot_uniform("encoder", 11, gain, br), Wt.glorot_uniform("encoder", 12, gain, br),
            Wt.glorot_uniform("decoder", 13, gain, br), Wt.glorot_uniform("decoder", 14, gain, br))


def make(name, img):
    out = {"input": img}
    H, W = img.shape[1:3]
    for wname in WEIGHT_SETS:
        eY, eC, dY, dC = weight_sets(wname)
        pre64 = O.encode_prequant(img, eY, eC, "f64")
        sym64 = O.quantise(pre64)
        sym32 = O.encode(img, eY, eC, "f32")
        rec64 = O.decode(sym64, dY, dC, "f64")
        rec32 = O.decode(sym64, dY, dC, "f32")
        hist, ent, bpp, hg = O.rate(sym64, H, W, "f32")
        # distance of every pre-round value from the nearest rounding tie, in symbol units (for the tie band)
        frac = pre64 * 255.0
        tie_dist = np.abs(frac - np.floor(frac) - 0.5).astype(np.float32)
        out.update({f"{wname}_sym64": sym64, f"{wname}_sym32": sym32, f"{wname}_rec64": rec64, f"{wname}_rec32": rec32,
                    f"{wname}_hist": hist.astype(np.uint32), f"{wname}_entropy": ent, f"{wname}_bpp": bpp,
                    f"{wname}_tie_dist": tie_dist})
        print(name, wname, "sym32 vs sym64 mismatch", np.mean(sym32 != sym64), "rec", np.mean(rec32 != rec64),
              "zeros", np.mean(sym64 == 0), "sat", np.mean(sym64 == 255), "bpp", bpp)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    kodim = np.array(Image.open(os.path.join(REF, "kodak_img", "kodim21.png")))
    assert kodim.shape == (512, 768, 3)
    Image.fromarray(kodim).save(os.path.join(OUT, "kodim21.png"), optimize=True)   # config-1 input for bench.py
    make("kodim21_crop", kodim[None, 192:320, 288:480])                          # 128 x 192 crop with the lighthouse
    patches = np.stack([np.array(Image.open(os.path.join(REF, "imagenet_patches", f"{i:05d}.jpg")).convert("RGB"))
                        for i in range(6)])
    assert patches.shape == (6, 128, 128, 3)
    make("imagenet_patches", patches)


if __name__ == "__main__":
    main()
