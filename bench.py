#!/usr/bin/env python
"""Benchmark of the codec hot path: encode -> rate estimate -> decode, megapixels per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl ours|reference]

One process per GPU (under torchrun for N > 1; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).
A step is one pass of the hot path over one synthetic batch that is already resident in HBM:
  c2 (default, BASELINE.json configs[1]): 24 x 512x768 RGB, encode + histogram/entropy rate + decode
  c1: one 768x512 Kodak image (tests/golden/kodim21.png), encode + rate + decode (launch-latency scale; L2 flushed between steps)
  c3: 4096 x 128x128 encode + rate      c4: 16 x 2160x3840 decode only
  c5: the FIXED set of 65 536 256x256 patches split over the ranks (strong scaling), encode + rate
With N GPUs every rank runs its own batch (weak scaling: images are independent; c5: its slice of the fixed set), the
symbol counts accumulate in a per-rank [3,256] device table and the ranks exchange ONE NCCL sum-allreduce of it per
run ("once per run / per macro-batch") -- inside the timed region, no per-step collective or barrier.
The default (c2) line also carries a `strong_c5` block: the batch-sharded configuration BASELINE.json names for the
1 -> 8 GPU scaling target, measured in the same process (value, e2e, all-rank histogram fingerprint).
Rank 0 prints ONE JSON line (see the field notes in DESIGN.md "Measurement").
--impl reference times the CPU restatement of the reference (oracle/, torch-CPU fp32, all host threads) on the same
workload and config keys; TensorFlow itself cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (images per GPU, H, W, stages, description)
    "c1": (1, 512, 768, ("encode", "rate", "decode"), "one Kodak 768x512 image (kodim21), encode+entropy estimate+decode"),
    "c2": (24, 512, 768, ("encode", "rate", "decode"), "Kodak-shape 24x768x512 full encode+entropy estimate+decode"),
    "c3": (4096, 128, 128, ("encode", "rate"), "ImageNet-patch 4096x128x128 encode + rate estimate"),
    "c4": (16, 2160, 3840, ("decode",), "3840x2160 x16 decode-only"),
    "c5": (65536, 256, 256, ("encode", "rate"), "65536 256x256 patches sharded over the GPUs, encode + histogram allreduce"),
}
# c5 is BASELINE.json's batch-sharded configuration: the 65536 patches are a FIXED set split over the ranks (strong scaling,
# SURVEY.md 8d/8e); the other workloads keep a fixed per-GPU batch (weak scaling).
STRONG = ("c5",)
WEIGHTS_NOTE = "random-init (Keras glorot-uniform, seeds 11-14)"

# algorithmic work per RGB pixel (3 colour planes), SURVEY.md 8a / 8d
FLOP_PER_PX = {"conv1": 1200, "conv2": 19200, "conv3": 13824, "conv4": 13824, "conv8": 4800, "dconv1": 4800,
               "dconv5": 13824, "dconv6": 13824, "dconv7": 38400, "dconv8": 2400}
BYTES_PER_PX = {"conv1": 3 + 96, "dconv8": 192 + 3, "hist": 1.5, "latent_expand": 1.5 + 6, "quantise": 7.5}
HBM_BOUND = ("conv1", "dconv8", "hist", "latent_expand", "quantise")
# The decoder's default path runs dconv8's tap-response GEMM inside dconv7 (which then writes 25 fp32 responses per output
# pixel instead of 64 split activations) and "dconv8" is the gather + colour + pack pass over those responses:
# 3 planes x 25 x 4 B / 4 final pixels = 75 B read + 3 B written per RGB pixel.  NNIC_FUSE_D78=0 restores the two-kernel form.
if os.environ.get("NNIC_FUSE_D78", "1") != "0":
    FLOP_PER_PX["dconv7"] += FLOP_PER_PX.pop("dconv8")
    BYTES_PER_PX["dconv8"] = 75 + 3


def measured_peaks(burst):
    """MEASURED_PEAKS.json (driver-written): copy GB/s and dense bf16 TFLOP/s.  A kernel timed inside a timed region
    shorter than about a second runs at burst clocks, so its denominator is the burst figure; a long run uses the
    sustained one.  The line says which."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        tf = p["bf16_tflops"] if burst else p.get("bf16_tflops_sustained", p["bf16_tflops"])
        return {"hbm_gbs": p["hbm_gbs"], "tflops": tf, "source": "measured", "which": "burst" if burst else "sustained"}
    return {"hbm_gbs": 6650.0, "tflops": 1630.0 if burst else 1400.0, "source": "fallback (B200_PROFILING.md)",
            "which": "burst" if burst else "sustained"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; started BEFORE the warm-up so that short timed regions
    still have samples under load (the first sample only arrives after nvidia-smi's own start-up)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Samples from here on count as 'in the timed region'."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        rows_all = [r for _t, r in self.rows]
        rows_timed = [r for t, r in self.rows if self.t_mark is not None and t >= self.t_mark]
        rows = rows_timed or rows_all            # a timed region shorter than one sampling period: warm-up samples (same load)
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "samples_in_timed_region": len(rows_timed), "samples_total": len(rows_all)}


def synthetic_batch_gpu(torch, n, h, w, seed, device):
    """Blocky + noisy uint8 RGB batch generated on the device (natural-ish statistics)."""
    g = torch.Generator(device=device).manual_seed(seed)
    base = torch.randint(0, 256, (n, (h + 7) // 8, (w + 7) // 8, 3), device=device, generator=g, dtype=torch.int16)
    up = base.repeat_interleave(8, dim=1).repeat_interleave(8, dim=2)[:, :h, :w].float()
    noise = torch.randn((n, h, w, 3), device=device, generator=g) * 10.0
    return (up + noise).clamp_(0, 255).to(torch.uint8)


def sharded_patches_gpu(torch, first, count, h, w, salt, device, chunk=1024):
    """uint8 RGB patches whose bytes depend only on (global patch index, position, salt), so every sharding of the patch
    set sees identical data and the all-rank symbol histogram must not depend on the number of GPUs.  Blocky (8x8) base
    plus uniform noise from an integer hash, generated on the device."""
    out = torch.empty((count, h, w, 3), dtype=torch.uint8, device=device)
    yy = torch.arange(h, device=device, dtype=torch.int64).view(1, h, 1, 1)
    xx = torch.arange(w, device=device, dtype=torch.int64).view(1, 1, w, 1)
    cc = torch.arange(3, device=device, dtype=torch.int64).view(1, 1, 1, 3)

    def mix(v):
        v = v & 0xFFFFFFFF
        v = ((v ^ (v >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
        v = ((v ^ (v >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
        return v ^ (v >> 16)
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        idx = (first + c0 + torch.arange(n, device=device, dtype=torch.int64)).view(n, 1, 1, 1)
        base = mix(idx * 0x9E3779B1 + (yy // 8) * 0x85EBCA6B + (xx // 8) * 0xC2B2AE35 + cc * 0x27D4EB2F + salt) & 255
        noise = mix(idx * 0x165667B1 + yy * 0xD3A2646C + xx * 0xFD7046C5 + cc * 0xB55A4F09 + salt + 1) % 41 - 20
        out[c0:c0 + n] = (base + noise).clamp_(0, 255).to(torch.uint8)
    return out


def synthetic_latent_gpu(torch, n, lh, lw, seed, device):
    """uint8 latents with a peaked (geometric) symbol distribution, like an encoder output."""
    g = torch.Generator(device=device).manual_seed(seed)
    u = torch.rand((n, lh, lw, 96), device=device, generator=g)
    return (torch.log1p(-u) / -0.08).clamp_(0, 255).to(torch.uint8)


def kodim21():
    import numpy as np
    from PIL import Image
    return np.array(Image.open(os.path.join(ROOT, "tests", "golden", "kodim21.png")))[None]


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference), all host threads
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, n_images, threads=None):
    """Oracle (fp32 restatement of the reference) on `n_images` of the workload; returns (MP/s, cores, text)."""
    import numpy as np
    import torch
    from neural_network_image_compression_b200 import weights as Wt
    from oracle import nnic_oracle as O
    torch.set_num_threads(threads or os.cpu_count() or 1)     # torchrun sets OMP_NUM_THREADS=1; use every host core
    cores = torch.get_num_threads()
    _n, H, W, stages, _d = WORKLOADS[workload]
    rng = np.random.default_rng(0)
    eY, eC = Wt.glorot_uniform("encoder", 11), Wt.glorot_uniform("encoder", 12)
    dY, dC = Wt.glorot_uniform("decoder", 13), Wt.glorot_uniform("decoder", 14)
    img = kodim21() if workload == "c1" else rng.integers(0, 256, size=(n_images, H, W, 3), dtype=np.uint8)
    lat = rng.integers(0, 64, size=(n_images, H // 8, W // 8, 96), dtype=np.uint8)
    t0 = time.perf_counter()
    chunk = 8 if H * W >= 512 * 768 else 256                 # bounded memory; the timing covers the whole sample
    for i0 in range(0, n_images, chunk):
        part = lat[i0:i0 + chunk]
        if "encode" in stages:
            part = O.encode(img[i0:i0 + chunk], eY, eC, "f32")
        if "rate" in stages:
            O.rate(part, H, W, "f32")
        if "decode" in stages:
            O.decode(part, dY, dC, "f32")
    dt = time.perf_counter() - t0
    return n_images * H * W / 1e6 / dt, cores, f"{n_images} x {H}x{W} images, {'+'.join(stages)}, torch-CPU fp32 oracle"


def base_config(workload, n_img):
    _n, H, W, stages, desc = WORKLOADS[workload]
    return {"workload": f"{workload}: {desc}", "images_per_gpu": n_img, "H": H, "W": W, "stages": list(stages),
            "weights": WEIGHTS_NOTE}


def run_reference(args):
    """Reference arm: the CPU restatement on the SAME per-GPU batch and config keys as our arm whenever K steps of it fit
    in a few minutes (c1, c2); larger workloads are timed on a bounded sample of the same shape, which `config` states."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img, H, W, stages, desc = WORKLOADS[args.workload]
    if args.workload in STRONG:
        n_img = n_img // max(1, args.gpus)
    est_mp_s = 1.3                                                      # torch-CPU fp32, 16 cores (profiles/)
    budget_s = 240.0
    full_s = (args.steps + 1) * n_img * H * W / 1e6 / est_mp_s
    sample = n_img if full_s <= budget_s else max(1, int(budget_s * est_mp_s * 1e6 / ((args.steps + 1) * H * W)))
    cpu_reference_sample(args.workload, min(sample, 2))                  # warm-up (thread pools, oneDNN primitives)
    vals, t0 = [], time.perf_counter()
    text, cores = "", 0
    for _ in range(args.steps):
        v, cores, text = cpu_reference_sample(args.workload, sample)
        vals.append(v)
    total = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    cfg = base_config(args.workload, n_img)
    cfg["sample_per_step"] = text if sample != n_img else "the whole per-GPU batch"
    cfg["l2"] = "n/a (CPU arm)"
    out = {"impl": "reference", "metric": "encode+decode megapixels/sec", "value": round(value, 4), "unit": "MP/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / args.steps * 1e3, 2),
           "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": round(value, 4), "unit": "MP/s", "cores": cores, "kind": "port", "sample": text},
           "e2e": {"value": round(value, 4), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "CPU restatement of the reference (oracle/), not TensorFlow: TensorFlow is not installable here"}
    emit(out)


_REAL_STDOUT = None
_ORIG_AFFINITY = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def pin_to_gpu_numa_node(local, world):
    """Host staging next to the GPU: bind this rank to the CPUs NVML reports for its GPU (its NUMA node), split among the
    ranks that share the same mask, before any pinned buffer is allocated (first touch places the pages)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = os.sched_getaffinity(0)
        avail = sorted(set(cpus) & _ORIG_AFFINITY)
        if not avail:
            return None
        if world > 1 and len(avail) >= 2 * world:
            per = len(avail) // world
            avail = avail[(local % world) * per:(local % world + 1) * per]
        os.sched_setaffinity(0, avail)
        return f"{len(avail)} CPUs ({avail[0]}-{avail[-1]})"
    except Exception as e:           # affinity is an optimisation, never a failure
        return f"unchanged ({type(e).__name__})"


# ---------------------------------------------------------------------------------------------------------------
# one workload on this rank
# ---------------------------------------------------------------------------------------------------------------
def measure(nn, torch, np, args, workload, steps, warmup, rank, world, local, profile, per_gpu_images=0, on_timed_start=None):
    dev = torch.device("cuda", local)
    n_img, H, W, stages, _desc = WORKLOADS[workload]
    strong = workload in STRONG and not per_gpu_images
    first_img = 0
    if strong:
        first_img, last_img = nn.dist.shard_range(n_img, rank, world)      # contiguous slice of the fixed set
        n_img = last_img - first_img
    if per_gpu_images:
        n_img = per_gpu_images
    lh, lw = H // 8, W // 8
    enc, dec = nn.Encoder(local, args.arith), nn.Decoder(local, args.arith)
    enc.init_random(); dec.init_random()

    in_bytes = n_img * H * W * 3 if "encode" in stages else n_img * lh * lw * 96
    flush = None
    if workload == "c1":
        # the input (1.2 MB) and every activation fit in L2: flush it between steps (a 256 MB write), timed regions are
        # the steps only (one event pair per step)
        n_sets = 1
        inputs = [torch.from_numpy(kodim21()).to(dev)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    elif strong:
        n_sets = 1 if in_bytes > 1 << 30 else 2       # a shard of >= 1 GB is far beyond L2 on its own
        inputs = [sharded_patches_gpu(torch, first_img, n_img, H, W, 7919 * i, dev) for i in range(n_sets)]
    elif "encode" in stages:
        n_sets = max(2, min(8, -(-160_000_000 // in_bytes)))
        inputs = [synthetic_batch_gpu(torch, n_img, H, W, 1000 * rank + i, dev) for i in range(n_sets)]
    else:
        n_sets = max(2, min(8, -(-160_000_000 // in_bytes)))
        inputs = [synthetic_latent_gpu(torch, n_img, lh, lw, 1000 * rank + i, dev) for i in range(n_sets)]
    lat_buf = torch.empty((n_img, lh, lw, 96), dtype=torch.uint8, device=dev)
    rgb_buf = torch.empty((n_img, H, W, 3), dtype=torch.uint8, device=dev) if "decode" in stages else None
    hist_global = torch.zeros((3, 256), dtype=torch.int64, device=dev)

    def step(i):
        x = inputs[i % n_sets]
        lat = x
        if "encode" in stages and "rate" in stages:
            lat, _r = enc.encode_rate(x, out=lat_buf, hist_global=hist_global)   # symbols counted where they are quantised
        elif "encode" in stages:
            lat = enc(x, out=lat_buf)
        elif "rate" in stages:
            nn.rate(enc.handle, lat, H, W, hist_global=hist_global)
        if "decode" in stages:
            dec(lat, out=rgb_buf)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    if profile:
        enc.handle.set_profiling(True); dec.handle.set_profiling(True)
        enc.handle.profile_collect(); dec.handle.profile_collect()
    l0 = enc.handle.launch_count + dec.handle.launch_count
    hist_global.zero_()
    barrier()
    if on_timed_start:
        on_timed_start()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(warmup + i)
        if "rate" in stages:
            nn.dist.allreduce_histogram(hist_global)      # the path's only exchange: once per run (no-op at N = 1)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.fill_(i & 0xff)
            evs[i][0].record()
            step(warmup + i)
            evs[i][1].record()
        nn.dist.allreduce_histogram(hist_global)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = enc.handle.launch_count + dec.handle.launch_count - l0
    prof = {}
    if profile:
        for hnd in (enc.handle, dec.handle):
            for k, (t, c) in hnd.profile_collect().items():
                prof[k] = (prof.get(k, (0.0, 0))[0] + t, prof.get(k, (0, 0))[1] + c)
        enc.handle.set_profiling(False); dec.handle.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    mp_per_step = (WORKLOADS[workload][0] if strong else world * n_img) * H * W / 1e6
    value = mp_per_step * steps / (ms / 1e3)

    # fingerprint of the all-rank symbol histogram of ONE pass over input set 0 (untimed): for the sharded workload it
    # must not depend on --gpus
    hist_sha = None
    if "rate" in stages and "encode" in stages:
        import hashlib
        hist_global.zero_()
        step(0)
        nn.dist.allreduce_histogram(hist_global)
        torch.cuda.synchronize()
        hist_sha = hashlib.sha1(hist_global.cpu().numpy().tobytes()).hexdigest()[:16]

    # ---- end to end through the public API with pinned HOST buffers (H2D + D2H inside the timed region) ----
    def pinned(shape):
        return torch.empty(shape, dtype=torch.uint8, pin_memory=True).numpy()
    # host staging holds at most 8192 images; a larger per-GPU batch goes through it slice by slice (every slice is copied
    # in and out, so the byte counts and the timing are those of the whole batch)
    n_e2e = min(n_img, 8192)
    slices = -(-n_img // n_e2e)
    h_in = pinned((n_e2e,) + tuple(inputs[0].shape[1:])); h_in[...] = inputs[0][:n_e2e].cpu().numpy()
    h_lat = pinned((n_e2e, lh, lw, 96))
    h_rgb = pinned((n_e2e, H, W, 3)) if "decode" in stages else None
    hg_host = np.zeros((3, 256), np.uint64)
    counts = SimpleNamespace(h2d=0, d2h=0)

    def e2e_step():
        counts.h2d = counts.d2h = 0
        for _s in range(slices):
            lat = h_in
            if "encode" in stages and "rate" in stages:
                lat, r = enc.encode_rate(h_in, out=h_lat, hist_global=hg_host); counts.h2d += h_in.nbytes + 6144
                counts.d2h += h_lat.nbytes + r.hist.nbytes + r.entropy_bits.nbytes + r.bpp.nbytes + 6144
            elif "encode" in stages:
                lat = enc(h_in, out=h_lat); counts.h2d += h_in.nbytes; counts.d2h += h_lat.nbytes
            elif "rate" in stages:
                r = nn.rate(enc.handle, lat, H, W, hist_global=hg_host)
                counts.h2d += lat.nbytes + 6144; counts.d2h += r.hist.nbytes + r.entropy_bits.nbytes + r.bpp.nbytes + 6144
            if "decode" in stages:
                dec(lat, out=h_rgb); counts.h2d += lat.nbytes; counts.d2h += h_rgb.nbytes

    e2e_steps = max(2, min(steps, 10))
    for _ in range(2 if slices == 1 else 1):
        e2e_step()
    hg_host[...] = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    if "rate" in stages:
        hg = torch.from_numpy(hg_host.astype(np.int64)).to(dev)
        nn.dist.allreduce_histogram(hg)                   # once per run, like the device-resident loop
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = mp_per_step * e2e_steps / e2e_s
    l2_note = ("L2 flushed (256 MB write) between steps; each step timed by its own event pair" if flush is not None else
               f"inputs rotate over {n_sets} distinct batch(es) ({n_sets * in_bytes / 1e6:.0f} MB > 126 MB L2); "
               "each step streams > 1 GB of activations")
    return SimpleNamespace(ms=ms, steps=steps, value=value, mp_per_step=mp_per_step, n_img=n_img, H=H, W=W, stages=stages,
                           strong=strong, prof=prof, launches=launches, hist_sha=hist_sha, e2e_value=e2e_value,
                           e2e_steps=e2e_steps, h2d=counts.h2d, d2h=counts.d2h, l2_note=l2_note, enc=enc, dec=dec,
                           inputs=inputs, lat_buf=lat_buf, rgb_buf=rgb_buf, n_sets=n_sets, barrier=barrier)


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)          # libraries that print to fd 1 (e.g. the NCCL version banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--arith", default="tc_split", choices=("tc_split", "simt_f32"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong-c5", action="store_true", help="skip the batch-sharded c5 block of the default line")
    ap.add_argument("--per-gpu-images", type=int, default=0, help="override the per-GPU batch (development)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import neural_network_image_compression_b200 as nn

    local_env = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    affinity = pin_to_gpu_numa_node(local_env, int(os.environ.get("WORLD_SIZE", "1")))
    rank, world, local = nn.dist.init_process_group_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    m = measure(nn, torch, np, args, args.workload, args.steps, warmup, rank, world, local, profile=True,
                per_gpu_images=args.per_gpu_images, on_timed_start=sampler.mark if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    n_img, H, W, stages = m.n_img, m.H, m.W, m.stages
    ms, value = m.ms, m.value

    # ---- informational: the same steps with the optional fp16 decoder arithmetic (NOT the headline: `value` and `e2e`
    #      above use the decoder that is as exact as the encoder) ----
    decode_fp16 = None
    if "decode" in stages and args.arith == "tc_split" and args.workload != "c1":
        dec16 = nn.Decoder(local, args.arith, precision="fp16")
        for i in range(2):
            dec16.set_weights(i, m.dec.weights[i])
        rgb16 = torch.empty_like(m.rgb_buf)

        def step16(i):
            x = m.inputs[i % m.n_sets]
            lat = x
            if "encode" in stages:
                lat = m.enc(x, out=m.lat_buf)
            dec16(lat, out=rgb16)
            return lat
        lat = step16(0)
        m.dec(lat, out=m.rgb_buf)
        torch.cuda.synchronize()
        diff = (rgb16.to(torch.int16) - m.rgb_buf.to(torch.int16)).abs()
        frac, dmax = float((diff != 0).float().mean().item()), int(diff.max().item())
        del diff
        for i in range(warmup):
            step16(i)
        m.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k16 = max(3, min(args.steps, 50))
        f0.record()
        for i in range(k16):
            step16(warmup + i)
        f1.record()
        m.barrier()
        ms16 = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ms16], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms16 = float(t.item())
        decode_fp16 = {"value": round(m.mp_per_step * k16 / (ms16 / 1e3), 2), "unit": "MP/s", "ms_per_step": round(ms16 / k16, 4),
                       "steps": k16, "stages": [s_ for s_ in stages if s_ != "rate"],
                       "differing_bytes_frac": round(frac, 5), "max_abs_byte_diff": dmax,
                       "note": "Decoder(precision='fp16'): one fp16 product per MAC; within BASELINE's 0.01 dB PSNR, not byte-identical"}
        del dec16, rgb16

    # ---- c1 only: the same step replayed as a CUDA graph (GraphCodec), L2 flushed between replays ----
    graph_block = None
    if args.workload == "c1":
        gc = nn.GraphCodec(m.enc, m.dec, 1, H, W)
        gc.x.copy_(m.inputs[0])
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for _ in range(3):
            gc.run()
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            evs[i][0].record(); gc.run(); evs[i][1].record()
        torch.cuda.synchronize()
        gms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        graph_block = {"ms_per_step": round(gms, 4), "value": round(H * W / 1e6 / (gms / 1e3), 2), "unit": "MP/s",
                       "api": "GraphCodec(enc, dec, 1, 512, 768).run(): encode_rate + decode captured once, replayed"}
        del gc, flush

    # ---- the batch-sharded configuration (BASELINE.json configs[4]) in the same process: strong scaling over the ranks ----
    strong_c5 = None
    if args.workload == "c2" and not args.no_strong_c5 and not args.per_gpu_images:
        keep = (m.enc, m.dec)
        m.inputs = m.lat_buf = m.rgb_buf = None              # release the c2 buffers before the 12.9 GB / world shard
        torch.cuda.empty_cache()
        s5 = measure(nn, torch, np, args, "c5", 3, 1, rank, world, local, profile=False)
        strong_c5 = {"metric": "encode+rate megapixels/sec, 65536 x 256x256 patches split over the ranks",
                     "value": round(s5.value, 2), "unit": "MP/s", "scaling": "strong", "steps": s5.steps,
                     "ms_per_step": round(s5.ms / s5.steps, 3), "patches_per_gpu": s5.n_img,
                     "e2e": {"value": round(s5.e2e_value, 2), "unit": "MP/s", "h2d_bytes_per_step": int(s5.h2d),
                             "d2h_bytes_per_step": int(s5.d2h), "steps": s5.e2e_steps},
                     "global_symbol_histogram_sha1": s5.hist_sha,
                     "note": "the same 65 536 patches for every --gpus (bytes depend on the global patch index only); the "
                             "fingerprint of the all-rank symbol counts must not change with the number of GPUs"}
        del s5, keep

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (event-timed inside the timed region) ----
    prof = m.prof
    peaks = measured_peaks(burst=ms < 1000.0)
    px_per_launch = n_img * H * W          # RGB pixels one launch of a kernel covers (micro-batches: see below)
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    roofline = None
    kernels = {}
    for k, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        launches_per_step = c / args.steps
        px = px_per_launch / launches_per_step            # pixels per launch
        avg_ms = t / c
        entry = {"ms_per_launch": round(avg_ms, 4), "launches_per_step": launches_per_step,
                 "share_of_step": round(t / ms, 4)}
        if k in HBM_BOUND and k in BYTES_PER_PX:
            entry["GB/s"] = round(BYTES_PER_PX[k] * px / avg_ms / 1e6, 1)
            entry["frac_of_hbm_peak"] = round(entry["GB/s"] / peaks["hbm_gbs"], 4)
        if k in FLOP_PER_PX and k not in HBM_BOUND:
            entry["TFLOP/s"] = round(FLOP_PER_PX[k] * px / avg_ms / 1e9, 1)
        kernels[k] = entry
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = {}
    if os.path.exists(traffic_path):
        with open(traffic_path) as f:
            traffic = json.load(f).get(args.workload, {})
    if dom:
        t, c = prof[dom]
        px = px_per_launch / (c / args.steps)
        avg_ms = t / c
        if dom in HBM_BOUND:
            ach = BYTES_PER_PX[dom] * px / avg_ms / 1e6
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic.get(dom), "peak_source": peaks["source"]}
        else:
            ach = FLOP_PER_PX[dom] * px / avg_ms / 1e9
            roofline = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": round(ach / peaks["tflops"], 4), "traffic": traffic.get(dom), "peak_source": peaks["source"],
                        "peak_kind": f"{peaks['which']} bf16 (timed region {ms / 1e3:.2f} s)",
                        "issued_tflops": round(3 * ach, 1), "issued_frac": round(3 * ach / peaks["tflops"], 4),
                        "note": "achieved/frac count ALGORITHMIC FLOPs; the fp16 hi/lo split issues three fp16 products per "
                                "MAC (issued_* = what the tensor pipe executes; ncu tensor-pipe activity in profiles/)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        if _ORIG_AFFINITY:
            os.sched_setaffinity(0, _ORIG_AFFINITY)                 # the CPU arm may use every host core
        sample = max(1, min(n_img, int(round(2.4e6 / (H * W))) or 1))
        cpu_reference_sample(args.workload, 1)                      # warm-up
        v, cores, text = cpu_reference_sample(args.workload, sample)
        cpu_baseline = {"value": round(v, 4), "unit": "MP/s", "cores": cores, "kind": "port", "sample": text}

    cfg = base_config(args.workload, n_img)
    cfg.update({**({"global_symbol_histogram_sha1": m.hist_sha} if m.hist_sha else {}), "arith": args.arith, "l2": m.l2_note,
                "exchange": "one NCCL sum-allreduce of the [3,256] symbol counts per run, inside the timed region",
                "host_affinity": affinity})
    out = {"metric": "encode+decode megapixels/sec", "value": round(value, 2), "unit": "MP/s", "n_gpus": world,
           "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
           "scaling": "strong" if m.strong else "weak", "vs_baseline": None,
           "dtype": "f16x2-split/f32-accumulate" if args.arith == "tc_split" else "f32",
           "data": "kodim21.png (tests/golden)" if args.workload == "c1" else "synthetic",
           "config": cfg,
           "e2e": {"value": round(m.e2e_value, 2), "unit": "MP/s", "h2d_bytes_per_step": int(m.h2d), "d2h_bytes_per_step": int(m.d2h),
                   "steps": m.e2e_steps, "api": "Encoder().encode_rate(x) / Decoder()(x) on pinned NumPy buffers"},
           "gpu_launches": int(m.launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
           "cpu_baseline": cpu_baseline}
    if decode_fp16:
        out["decode_fp16"] = decode_fp16
    if graph_block:
        out["cuda_graph"] = graph_block
    if strong_c5:
        out["strong_c5"] = strong_c5
    emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
