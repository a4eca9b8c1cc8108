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


def test_c_abi_hist_allreduce_over_nccl(nn):
    """nnic_hist_allreduce (include/nnic.h): the exchange step through the C ABI with the caller's ncclComm_t.  One
    process drives every visible GPU (up to 2) with communicators from ncclCommInitAll; on a one-GPU box the
    communicator has one rank and the sum is the identity."""
    import ctypes as C
    import torch
    ndev = min(torch.cuda.device_count(), 2)
    nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)
    comms = (C.c_void_p * ndev)()
    devs = (C.c_int * ndev)(*range(ndev))
    assert nccl.ncclCommInitAll(comms, ndev, devs) == 0
    try:
        rng = np.random.default_rng(2)
        lats = [np.minimum(rng.geometric(0.15, size=(3, 8, 12, 96)) - 1, 255).astype(np.uint8) for _ in range(ndev)]
        handles = [nn.Handle(d) for d in range(ndev)]
        hgs = []
        for d in range(ndev):
            with torch.cuda.device(d):
                r = nn.rate(handles[d], torch.from_numpy(lats[d]).cuda(d), 64, 96)
                hgs.append(r.hist_global)
                torch.cuda.synchronize(d)
        assert nccl.ncclGroupStart() == 0
        for d in range(ndev):
            h = handles[d]
            h.check(h.lib.nnic_hist_allreduce(h.h, comms[d], hgs[d].data_ptr(), None), "nnic_hist_allreduce")
        assert nccl.ncclGroupEnd() == 0
        want = sum(O.histogram(lat).sum(axis=0) for lat in lats)
        for d in range(ndev):
            torch.cuda.synchronize(d)
            assert np.array_equal(hgs[d].cpu().numpy(), want)
        with pytest.raises(nn.NnicError):
            handles[0].check(handles[0].lib.nnic_hist_allreduce(handles[0].h, None, hgs[0].data_ptr(), None), "null comm")
    finally:
        for d in range(ndev):
            nccl.ncclCommDestroy(C.c_void_p(comms[d]))


def test_one_process_drives_two_gpus(nn):
    """One handle per GPU inside ONE process (SURVEY.md 8b threading row): the same weights and images give the same
    bytes on every device; kernel attributes and tensor maps are set up per device."""
    import torch
    from conftest import make_weights, synthetic_images
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    eY, eC, dY, dC = make_weights("spread")
    img = synthetic_images(3, 72, 104, seed=91)
    outs = []
    for dev in (1, 0, 1):
        enc, dec = nn.Encoder(dev), nn.Decoder(dev)
        enc.set_weights(0, eY); enc.set_weights(1, eC); dec.set_weights(0, dY); dec.set_weights(1, dC)
        lat, r = enc.encode_rate(img)
        x = torch.from_numpy(img).cuda(dev)
        lat_d = enc(x)
        assert lat_d.device.index == dev and np.array_equal(lat_d.cpu().numpy(), lat)
        outs.append((lat, r.hist.copy(), dec(lat)))
    for lat, hist, rec in outs[1:]:
        assert np.array_equal(lat, outs[0][0]) and np.array_equal(hist, outs[0][1]) and np.array_equal(rec, outs[0][2])
    assert np.array_equal(outs[0][1].astype(np.int64), O.histogram(outs[0][0]))
    enc0 = nn.Encoder(0)
    enc0.set_weights(0, eY); enc0.set_weights(1, eC)
    with pytest.raises(ValueError, match="handle's GPU"):
        enc0(torch.from_numpy(img).cuda(1))
    with pytest.raises(ValueError, match="handle's GPU"):
        nn.rate(enc0.handle, torch.from_numpy(outs[0][0]).cuda(1))
