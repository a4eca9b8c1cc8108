"""Device-resident encode + rate + decode of a small batch, launched directly and replayed as a CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_network_image_compression_b200 as nn


def main():
    n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 512, 768)))
    enc, dec = nn.Encoder(0).init_random(), nn.Decoder(0).init_random()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    lat = torch.empty((n, h // 8, w // 8, 96), dtype=torch.uint8, device="cuda")
    rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    hg = torch.zeros((3, 256), dtype=torch.int64, device="cuda")

    def step():
        hg.zero_()
        _, r = enc.encode_rate(x, out=lat, hist_global=hg)
        dec(lat, out=rgb)
        return r

    def timed(fn, reps=300):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()                       # scratch buffers reach their final size before the capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref_lat, ref_rgb = lat.clone(), rgb.clone()
    t_direct = timed(step)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        r = step()
    lat.zero_(); rgb.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(lat, ref_lat) and torch.equal(rgb, ref_rgb) and int(hg.sum()) == lat.numel()
    t_graph = timed(graph.replay)
    mp = n * h * w / 1e6
    print(f"{n}x{h}x{w}: direct {t_direct * 1e3:.1f} us ({mp / t_direct * 1e3:.0f} MP/s), CUDA graph {t_graph * 1e3:.1f} us "
          f"({mp / t_graph * 1e3:.0f} MP/s), bpp[0] {float(r.bpp[0]):.4f}")


if __name__ == "__main__":
    main()
