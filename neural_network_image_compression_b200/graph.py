"""CUDA-graph fast path for repeated shapes (small calls are launch-bound: one 768x512 image is 13 kernels in ~0.25 ms).

With device buffers the library only enqueues kernels and memsets on the caller's stream, and its scratch addresses and
tensor maps repeat from call to call, so encode_rate + decode of a fixed shape can be captured once and replayed:
`GraphCodec(enc, dec, n, H, W)` owns static device buffers, `run(x)` copies the batch in (device to device, or pinned
host to device), replays the graph and returns views of the static outputs.  Results are bit-identical to the direct
calls (tests/test_gpu_parity.py::test_graph_codec_matches_direct_calls)."""
from __future__ import annotations


class GraphCodec:
    def __init__(self, enc, dec=None, n: int = 1, H: int = 512, W: int = 768, warmup: int = 2):
        import torch
        if dec is not None and dec.device != enc.device:
            raise ValueError("encoder and decoder must share a GPU")
        dev = torch.device("cuda", enc.device)
        lh, lw = -(-H // 8), -(-W // 8)
        self.enc, self.dec, self.shape = enc, dec, (n, H, W)
        self.x = torch.zeros((n, H, W, 3), dtype=torch.uint8, device=dev)
        self.latent = torch.empty((n, lh, lw, 96), dtype=torch.uint8, device=dev)
        self.rgb = torch.empty((n, 8 * lh, 8 * lw, 3), dtype=torch.uint8, device=dev) if dec is not None else None
        self.hist_global = torch.zeros((3, 256), dtype=torch.int64, device=dev)
        self.rate = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # scratch buffers reach their final size before the capture
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.rate = self._step()

    def _step(self):
        self.hist_global.zero_()
        _lat, r = self.enc.encode_rate(self.x, out=self.latent, hist_global=self.hist_global)
        if self.dec is not None:
            self.dec(self.latent, out=self.rgb)
        return r

    def run(self, x=None):
        """x: uint8 [n,H,W,3] CUDA tensor, pinned/pageable CPU tensor or NumPy array (None: reuse the static input).
        Returns (latent, Rate, rgb) -- views of the static device buffers, valid until the next run()."""
        if x is not None:
            import numpy as np
            import torch
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(x)
            if tuple(x.shape) != tuple(self.x.shape) or x.dtype != torch.uint8:
                raise ValueError(f"expected uint8 {tuple(self.x.shape)}")
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.latent, self.rate, self.rgb
