"""GPU side of the multi-GPU path: NCCL allreduce of the CUDA histogram tensor (a 1-rank group on a 1-GPU
box; bench.py --gpus N drives the same functions with N ranks) and the global rate from reduced counts."""
import os
import socket

import numpy as np
import pytest

from oracle import nnic_oracle as O

pytestmark = pytest.mark.gpu


def test_nccl_histogram_allreduce_and_global_rate(nn, codec_factory):
    import torch
    import torch.distributed as dist
    enc, _ = codec_factory("default", "tc_split")
    rng = np.random.default_rng(1)
    lat = np.minimum(rng.geometric(0.2, size=(4, 16, 24, 96)) - 1, 255).astype(np.uint8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        lo, hi = nn.dist.shard_range(lat.shape[0], 0, 1)
        r = nn.rate(enc.handle, torch.from_numpy(lat[lo:hi]).cuda(), 128, 192)
        nn.dist.allreduce_histogram(r.hist_global)
        want = O.histogram(lat).sum(axis=0)
        assert np.array_equal(r.hist_global.cpu().numpy(), want)
        e, bpp = nn.dist.global_rate(enc.handle, r.hist_global, 16, 24, 128, 192)
        assert np.abs(e - O.entropy_from_hist(want)).max() < 1e-5
    finally:
        dist.destroy_process_group()
