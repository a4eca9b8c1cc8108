import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import neural_network_image_compression_b200 as nn
import bench
enc, dec = nn.Encoder(0).init_random(), nn.Decoder(0).init_random()
x = bench.synthetic_batch_gpu(torch, 24, 512, 768, 1, torch.device('cuda')).cpu()
h_in = torch.empty((24,512,768,3), dtype=torch.uint8, pin_memory=True); h_in.copy_(x); h_in = h_in.numpy()
h_lat = torch.empty((24,64,96,96), dtype=torch.uint8, pin_memory=True).numpy()
h_rgb = torch.empty((24,512,768,3), dtype=torch.uint8, pin_memory=True).numpy()
for rep in range(2):
  for mb in (0, 8, 6, 4, 3, 2):
    enc.handle.set_micro_batch(mb); dec.handle.set_micro_batch(mb)
    for _ in range(3):
        enc.encode_rate(h_in, out=h_lat); dec(h_lat, out=h_rgb)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        enc.encode_rate(h_in, out=h_lat); dec(h_lat, out=h_rgb)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"micro_batch {mb}: {dt*1e3:.3f} ms/step  {24*512*768/1e6/dt:.0f} MP/s", flush=True)
