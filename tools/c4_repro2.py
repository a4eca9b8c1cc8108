"""Development aid: the device-path part of bench.measure for c4 with switches: python tools/c4_repro2.py enc=1 sets=2 prof=1 zero=1 events=1 warm=3 steps=3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import neural_network_image_compression_b200 as nn

o = dict(enc=1, sets=2, prof=1, zero=1, events=1, warm=3, steps=3, n=16, lh=270, lw=480, sync=0, same_data=0, same_ptr=0)
for a in sys.argv[1:]:
    k, v = a.split("=")
    o[k] = int(v)
dev = torch.device("cuda", 0)
if o["enc"]:
    enc = nn.Encoder(0); enc.init_random()
dec = nn.Decoder(0); dec.init_random()
inputs = [bench.synthetic_latent_gpu(torch, o["n"], o["lh"], o["lw"], i, dev) for i in range(o["sets"])]
if o["same_data"]:                       # two tensors (two addresses), identical bytes
    inputs = [inputs[0], inputs[0].clone()] if o["sets"] == 2 else inputs
stage = torch.empty_like(inputs[0])


def pick(i):
    x = inputs[i % o["sets"]]
    if o["same_ptr"]:                    # one address, alternating bytes
        stage.copy_(x)
        return stage
    return x


rgb = torch.empty((o["n"], 8 * o["lh"], 8 * o["lw"], 3), dtype=torch.uint8, device=dev)
hg = torch.zeros((3, 256), dtype=torch.int64, device=dev)
try:
    for i in range(o["warm"]):
        dec(pick(i), out=rgb)
    torch.cuda.synchronize()
    if o["prof"]:
        dec.handle.set_profiling(True); dec.handle.profile_collect()
    if o["zero"]:
        hg.zero_()
    torch.cuda.synchronize()
    if o["events"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for i in range(o["steps"]):
        dec(pick(o["warm"] + i), out=rgb)
        if o["sync"]:
            torch.cuda.synchronize()
    if o["events"]:
        e1.record()
    torch.cuda.synchronize()
    print("ok", " ".join(sys.argv[1:]))
except Exception as e:                                  # noqa: BLE001
    print("FAILED", " ".join(sys.argv[1:]), "|", str(e)[:120].replace("\n", " "))
