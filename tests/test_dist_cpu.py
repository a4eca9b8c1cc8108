"""world_size-2 gloo test of the multi-GPU host logic: batch sharding + histogram allreduce.
The per-rank histograms come from the oracle here (no GPU); on the GPU box the same functions are
driven with NCCL tensors by bench.py and tests/test_gpu_dist.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, latent, out_dir):
    sys.path.insert(0, ROOT)
    import neural_network_image_compression_b200 as nn
    from oracle import nnic_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = nn.dist.init_process_group_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = nn.dist.shard_range(latent.shape[0], rank, world)
    local = O.histogram(latent[lo:hi]).sum(axis=0)                 # this rank's [3,256] counts
    t = torch.from_numpy(local.astype(np.int64))
    nn.dist.allreduce_histogram(t)
    arr = local.astype(np.uint64)
    nn.dist.allreduce_histogram(arr)                               # NumPy path
    assert np.array_equal(arr.astype(np.int64), t.numpy())
    np.save(os.path.join(out_dir, f"hist_{rank}.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_histogram_allreduce_world2(tmp_path):
    from oracle import nnic_oracle as O
    rng = np.random.default_rng(5)
    latent = rng.integers(0, 256, size=(5, 4, 6, 96), dtype=np.uint8)   # 5 images: ragged split 3 + 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, latent, str(tmp_path)), nprocs=2, join=True)
    want = O.histogram(latent).sum(axis=0)
    for r in range(2):
        got = np.load(tmp_path / f"hist_{r}.npy")
        assert np.array_equal(got, want)          # integer sums: bit-identical on every rank


def test_allreduce_is_noop_without_process_group():
    import neural_network_image_compression_b200 as nn
    h = np.arange(768, dtype=np.uint64).reshape(3, 256)
    assert nn.dist.allreduce_histogram(h.copy()).tolist() == h.tolist()
