"""Quick device-resident timing of encode / rate / decode (CUDA events) -- development aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_network_image_compression_b200 as nn


def main():
    N, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (24, 512, 768)))
    arith = sys.argv[4] if len(sys.argv) > 4 else "tc_split"
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    mb = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    enc, dec = nn.Encoder(0, arith), nn.Decoder(0, arith)
    enc.init_random(); dec.init_random()
    enc.handle.set_micro_batch(mb); dec.handle.set_micro_batch(mb)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tot = []
    for it in range(reps + 2):
        ev[0].record()
        lat = enc(x)
        ev[1].record()
        r = nn.rate(enc.handle, lat, H, W)
        ev[2].record()
        rec = dec(lat)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 2 and reps > 20:
            tot.append(ev[0].elapsed_time(ev[3]))
        elif it >= 2:
            te, tr, td = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
            mp = N * H * W / 1e6
            print(f"{arith} {N}x{H}x{W} mb={mb}: encode {te:.3f} ms ({mp / te * 1e3:.0f} MP/s)  rate {tr:.3f} ms  "
                  f"decode {td:.3f} ms ({mp / td * 1e3:.0f} MP/s)  enc+rate+dec {mp / (te + tr + td) * 1e3:.0f} MP/s", flush=True)
    if tot:
        half = tot[len(tot) // 2:]
        ms = sum(half) / len(half)
        print(f"{os.environ.get('NNIC_LIB', 'libnnic.so').split('/')[-1]} {arith} {N}x{H}x{W}: sustained {ms:.4f} ms/step "
              f"({N * H * W / 1e3 / ms:.0f} MP/s) over the last {len(half)} of {reps} steps", flush=True)


if __name__ == "__main__":
    main()
