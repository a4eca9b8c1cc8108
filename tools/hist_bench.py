"""Standalone histogram pass (nnic_rate on a latent that is already in HBM) at the config-5 scale: GB/s per k_hist variant.
NNIC_HIST_VARIANT = copies * 100 + resident blocks per SM (development switch read at nnic_create)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_network_image_compression_b200 as nn


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    variants = [int(v) for v in sys.argv[2:]] or [0]
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand((n, 32, 32, 96), device="cuda", generator=g)
    lats = {"geometric(0.08), 8% zeros": (torch.log1p(-u) / -0.08).clamp_(0, 255).to(torch.uint8),
            "encoder-like, 50% zeros, 50 symbols": torch.where(u < 0.5, torch.zeros_like(u), torch.log1p(-(u - 0.5) * 2) / -0.12 + 1).clamp_(0, 49).to(torch.uint8),
            "uniform 0..255": (u * 256).clamp_(0, 255).to(torch.uint8)}
    del u
    want = {k: torch.stack([torch.bincount(v[..., 32 * p:32 * p + 32].reshape(-1).int(), minlength=256) for p in range(3)]) for k, v in lats.items()}
    for var in variants:
        os.environ["NNIC_HIST_VARIANT"] = str(var)
        h = nn.Handle(0)
        for name, lat in lats.items():
            hg = torch.zeros((3, 256), dtype=torch.int64, device="cuda")
            r = nn.rate(h, lat, 256, 256, hist_global=hg)
            torch.cuda.synchronize()
            ok = bool(torch.equal(hg, want[name])) and int(r.hist.sum()) == lat.numel()
            h.set_profiling(True); h.profile_collect()
            for _ in range(5):
                nn.rate(h, lat, 256, 256)
            torch.cuda.synchronize()
            t, c = h.profile_collect()["hist"]
            h.set_profiling(False)
            print(f"variant {var:5d}  {name:38s} {lat.numel() / 1e6:7.1f} MB  {t / c:7.4f} ms  {lat.numel() / (t / c) / 1e6:7.1f} GB/s  "
                  f"({lat.numel() / (t / c) / 1e6 / 6528.4 * 100:4.1f}% of the copy peak)  counts {'ok' if ok else 'WRONG'}", flush=True)
        h.close()


if __name__ == "__main__":
    main()
