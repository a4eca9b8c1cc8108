"""Stress loop of the device-buffer decoder (development aid): python tools/stress_decode.py N lh lw iterations
On a failure the next call reports the barrier-wait code a timed-out tensor-core kernel left in the handle's host flag."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_network_image_compression_b200 as nn


def main():
    n, lh, lw, iters = (int(v) for v in sys.argv[1:5])
    profile = len(sys.argv) > 5 and sys.argv[5] == "profile"     # per-launch events, no synchronisation between iterations
    dec = nn.Decoder(0)
    dec.init_random()
    if profile:
        dec.handle.set_profiling(True)
    g = torch.Generator(device="cuda").manual_seed(1)
    mode = sys.argv[6] if len(sys.argv) > 6 else "uniform"
    if mode == "geometric":                      # bench.py's latents: peaked symbol distribution
        u = torch.rand((n, lh, lw, 96), device="cuda", generator=g)
        lat = (torch.log1p(-u) / -0.08).clamp_(0, 255).to(torch.uint8)
    else:
        lat = torch.randint(0, 256, (n, lh, lw, 96), dtype=torch.uint8, device="cuda", generator=g)
    if len(sys.argv) > 7 and sys.argv[7] == "enc":
        enc = nn.Encoder(0); enc.init_random()
    out = torch.empty((n, 8 * lh, 8 * lw, 3), dtype=torch.uint8, device="cuda")
    ref = None
    t0 = time.time()
    for i in range(iters):
        try:
            dec(lat, out=out)
            if not profile or i == 0 or i == iters - 1:
                torch.cuda.synchronize()
        except Exception as e:                      # noqa: BLE001
            print(f"iteration {i}: {type(e).__name__}: {str(e)[:300]}", flush=True)
            try:
                dec(lat, out=out)
            except Exception as e2:                 # noqa: BLE001
                print(f"next call: {str(e2)[:300]}", flush=True)
            return 1
        if ref is None:
            ref = out.clone()
        elif not torch.equal(ref, out):
            d = (ref != out)
            idx = d.nonzero()
            lo, hi = idx.min(dim=0).values.tolist(), idx.max(dim=0).values.tolist()
            per_ch = d.sum(dim=(0, 1, 2)).tolist()
            maxabs = (ref.to(torch.int16) - out.to(torch.int16)).abs().max().item()
            print(f"iteration {i}: output differs from iteration 0 in {d.sum().item()} bytes: images {lo[0]}..{hi[0]}, rows {lo[1]}..{hi[1]}, "
                  f"columns {lo[2]}..{hi[2]}, per RGB channel {per_ch}, max |difference| {maxabs}", flush=True)
            out2 = torch.empty_like(out)
            dec(lat, out=out2)
            torch.cuda.synchronize()
            print(f"   repeated once more: equals iteration 0: {torch.equal(out2, ref)}, equals the deviating output: {torch.equal(out2, out)}", flush=True)
            return 2
    print(f"{iters} iterations of {n}x{8 * lh}x{8 * lw} ok, identical bytes, {time.time() - t0:.1f} s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
