"""Pure-write, pure-read and copy bandwidth of HBM on this GPU (torch kernels, CUDA events): the denominator question for kernels
whose algorithmic bytes are almost all WRITES (conv1: 3 B read + 96 B written per pixel)."""
import torch

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n = 2 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)
t = timed(lambda: a.zero_());            print(f"write (memset, 2 GiB):        {n / t / 1e6:8.1f} GB/s")
t = timed(lambda: af.fill_(1.5));        print(f"write (fill kernel, 2 GiB):   {n / t / 1e6:8.1f} GB/s")
t = timed(lambda: af.sum());             print(f"read  (sum reduction, 2 GiB): {n / t / 1e6:8.1f} GB/s")
t = timed(lambda: b.copy_(a));           print(f"copy  (2 GiB + 2 GiB):        {2 * n / t / 1e6:8.1f} GB/s read+write")
