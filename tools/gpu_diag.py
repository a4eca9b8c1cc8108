"""GPU diagnostic: layer-by-layer comparison of both arithmetic modes of libnnic.so against the oracle.
Run on a B200 (`gpurun -- python tools/gpu_diag.py`).  Test infrastructure: imports oracle/."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_network_image_compression_b200 as nn
from neural_network_image_compression_b200 import weights as Wt
from oracle import nnic_oracle as O


def stats(name, got, ref):
    got = np.asarray(got, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    if got.shape != ref.shape:
        print(f"  {name}: SHAPE MISMATCH {got.shape} vs {ref.shape}")
        return
    d = np.abs(got - ref)
    scale = np.abs(ref).max() + 1e-30
    print(f"  {name}: max|d|={d.max():.3e} rel-to-max={d.max() / scale:.3e} mean|d|={d.mean():.3e} "
          f"ref-max={scale:.3e} nan={np.isnan(got).sum()}")


def main():
    arith_list = sys.argv[1].split(",") if len(sys.argv) > 1 else ["simt_f32", "tc_split"]
    rng = np.random.default_rng(0)
    N, H, W = 2, 64, 96
    # smooth-ish random image
    base = rng.integers(0, 256, size=(N, H // 4, W // 4, 3)).astype(np.float32)
    img = np.clip(np.kron(base, np.ones((1, 4, 4, 1))) + rng.normal(0, 12, size=(N, H, W, 3)), 0, 255).astype(np.uint8)
    gain, br = 1.6, 0.05
    wsets = {"encY": Wt.glorot_uniform("encoder", 11, gain, br), "encCbCr": Wt.glorot_uniform("encoder", 12, gain, br),
             "decY": Wt.glorot_uniform("decoder", 13, gain, br), "decCbCr": Wt.glorot_uniform("decoder", 14, gain, br)}
    planes = O.rgb_to_planes(img, "f32")
    tr = [O.encoder_trace(planes[0], wsets["encY"], "f64"), O.encoder_trace(planes[1], wsets["encCbCr"], "f64"),
          O.encoder_trace(planes[2], wsets["encCbCr"], "f64")]
    ref_layers = [np.concatenate([tr[p][i] for p in range(3)], axis=0) for i in range(5)]   # plane-major batches
    sym_ref = O.encode(img, wsets["encY"], wsets["encCbCr"], "f32")
    sym_ref64 = O.encode(img, wsets["encY"], wsets["encCbCr"], "f64")
    lat_planes = [(sym_ref.astype(np.float32) / np.float32(255))[..., 32 * i:32 * (i + 1)] for i in range(3)]
    dtr = [O.decoder_trace(lat_planes[0], wsets["decY"], "f64"), O.decoder_trace(lat_planes[1], wsets["decCbCr"], "f64"),
           O.decoder_trace(lat_planes[2], wsets["decCbCr"], "f64")]
    dref_layers = [np.concatenate([dtr[p][i] for p in range(3)], axis=0) for i in range(6)]
    rec_ref = O.decode(sym_ref, wsets["decY"], wsets["decCbCr"], "f32")
    rc = 0
    for arith in arith_list:
        print(f"=== arith {arith} ===", flush=True)
        t0 = time.time()
        enc = nn.Encoder(0, arith)
        enc.set_weights(0, wsets["encY"]); enc.set_weights(1, wsets["encCbCr"])
        dec = nn.Decoder(0, arith)
        dec.set_weights(0, wsets["decY"]); dec.set_weights(1, wsets["decCbCr"])
        enc.handle.set_micro_batch(N); dec.handle.set_micro_batch(N)   # one micro-batch, so debug_fetch sees every image
        sym, pre = enc(img, return_prequant=True)
        print(f"  encode done in {time.time() - t0:.2f}s; launches {enc.handle.launch_count}", flush=True)
        for slot, nm in enumerate(["conv1", "conv2", "conv3", "conv4+res"]):
            stats(nm, enc.handle.debug_fetch(slot), ref_layers[slot])
        stats("prequant", pre, O.encode_prequant(img, wsets["encY"], wsets["encCbCr"], "f64"))
        mm = np.mean(sym != sym_ref)
        mm64 = np.mean(sym != sym_ref64)
        print(f"  symbols: mismatch vs f32 oracle {mm:.3e}, vs f64 oracle {mm64:.3e}, max|diff| "
              f"{np.abs(sym.astype(int) - sym_ref.astype(int)).max()}", flush=True)
        rec, rpre = dec(sym_ref, return_prequant=True)
        for slot, nm in zip(range(4, 8), ["latent/255", "dconv1", "dconv5", "dconv6+res"]):
            stats(nm, dec.handle.debug_fetch(slot), dref_layers[slot - 4])
        outp = dec.run_model(lat_planes)
        stats("decoder planes", np.concatenate(outp, axis=0), dref_layers[5])
        rm = np.mean(rec != rec_ref)
        print(f"  recon: mismatch vs f32 oracle {rm:.3e}, max|diff| {np.abs(rec.astype(int) - rec_ref.astype(int)).max()}, "
              f"psnr(ref,got) {O.psnr(rec_ref, rec):.2f}", flush=True)
        r = nn.rate(enc.handle, sym_ref)
        hist, ent, bpp, hg = O.rate(sym_ref, H, W)
        print(f"  rate: hist equal {np.array_equal(r.hist.astype(np.int64), hist)}, global equal "
              f"{np.array_equal(r.hist_global.astype(np.int64), hg)}, max|dH| {np.abs(r.entropy_bits - ent).max():.2e}, "
              f"max|dbpp| {np.abs(r.bpp - bpp).max():.2e}", flush=True)
        if mm > 1e-3 or rm > 1e-3:
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
