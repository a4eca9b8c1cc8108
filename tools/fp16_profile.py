import sys, torch
sys.path.insert(0, ".")
import neural_network_image_compression_b200 as nn
for prec in ("split", "fp16"):
    dec = nn.Decoder(0, precision=prec).init_random()
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randint(0, 64, (24, 64, 96, 96), dtype=torch.uint8, device="cuda", generator=g)
    out = torch.empty((24, 512, 768, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3): dec(lat, out=out)
    torch.cuda.synchronize()
    dec.handle.set_profiling(True); dec.handle.profile_collect()
    for _ in range(50): dec(lat, out=out)
    torch.cuda.synchronize()
    p = dec.handle.profile_collect()
    print(prec, {k: round(t / c, 4) for k, (t, c) in p.items() if c})
