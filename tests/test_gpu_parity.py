"""GPU parity tests (run on a B200 with -m gpu): the CUDA path, called through the C ABI, against the oracle.

Tolerances are BASELINE.json's: quantised symbols bit-exact except at rounding ties (mismatch fraction
<= 1e-4, every mismatch is +-1 and lies in the documented tie band |v*255 - (k+1/2)| < 2e-3 of the fp64
oracle), histograms bit-exact, reconstructions within 0.01 dB PSNR and 1e-3 bpp.
"""
import numpy as np
import pytest

from conftest import load_golden, make_weights, synthetic_images
from oracle import nnic_oracle as O

pytestmark = pytest.mark.gpu

SYMBOL_MISMATCH_LIMIT = 1e-4
TIE_BAND = 2e-3          # symbol units
PSNR_TOL_DB = 0.01
BPP_TOL = 1e-3


def check_symbols(got, want, tie_dist=None, limit=SYMBOL_MISMATCH_LIMIT, min_allow=2):
    assert got.shape == want.shape and got.dtype == np.uint8
    diff = got != want
    n_bad = int(diff.sum())
    assert n_bad <= max(min_allow, limit * got.size), f"{n_bad} of {got.size} symbols differ"
    if n_bad:
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
        if tie_dist is not None:
            assert np.all(tie_dist[diff] < TIE_BAND), "a mismatch lies outside the rounding-tie band"
    return n_bad


@pytest.mark.parametrize("arith", ["tc_split", "simt_f32"])
@pytest.mark.parametrize("wname", ["default", "spread"])
@pytest.mark.parametrize("name", ["kodim21_crop", "imagenet_patches"])
def test_golden_encode_rate_decode(nn, codec_factory, name, wname, arith):
    g = load_golden(name)
    enc, dec = codec_factory(wname, arith)
    img = g["input"]
    H, W = img.shape[1:3]
    sym = enc(img)
    check_symbols(sym, g[f"{wname}_sym64"], g[f"{wname}_tie_dist"])
    check_symbols(sym, g[f"{wname}_sym32"])
    # rate of the golden latent: integer histogram bit-exact, entropy / bpp to fp32 rounding
    r = nn.rate(enc.handle, g[f"{wname}_sym64"], H, W)
    assert np.array_equal(r.hist, g[f"{wname}_hist"])
    assert np.array_equal(r.hist_global.astype(np.int64), g[f"{wname}_hist"].astype(np.int64).sum(axis=0))
    assert np.abs(r.entropy_bits - g[f"{wname}_entropy"]).max() < 1e-5
    assert np.abs(r.bpp - g[f"{wname}_bpp"]).max() < 1e-5
    # bpp of our own latent vs the reference latent
    assert np.abs(nn.rate(enc.handle, sym, H, W).bpp - g[f"{wname}_bpp"]).max() < BPP_TOL
    # decode the golden latent
    rec = dec(g[f"{wname}_sym64"])
    check_symbols(rec, g[f"{wname}_rec64"])
    check_symbols(rec, g[f"{wname}_rec32"])
    for i in range(img.shape[0]):
        assert abs(O.psnr(img[i], rec[i]) - O.psnr(img[i], g[f"{wname}_rec64"][i])) < PSNR_TOL_DB


@pytest.mark.parametrize("wname", ["default", "spread"])
def test_config1_kodim21_full(nn, codec_factory, wname):
    """BASELINE.json config 1: the reference's own CPU-runnable case, one Kodak 768x512 image."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    img = np.array(Image.open(os.path.join(GOLDEN, "kodim21.png")))[None]
    assert img.shape == (1, 512, 768, 3)
    eY, eC, dY, dC = make_weights(wname)
    enc, dec = codec_factory(wname, "tc_split")
    sym, pre = enc(img, return_prequant=True)
    pre64 = O.encode_prequant(img, eY, eC, "f64")
    sym64 = O.quantise(pre64)
    tie = np.abs(pre64 * 255.0 - np.floor(pre64 * 255.0) - 0.5)
    check_symbols(sym, sym64, tie)
    check_symbols(sym, O.encode(img, eY, eC, "f32"))
    assert np.abs(pre - pre64).max() < 2e-5
    rec = dec(sym64)
    rec_ref = O.decode(sym64, dY, dC, "f32")
    check_symbols(rec, rec_ref)
    assert abs(O.psnr(img, rec) - O.psnr(img, rec_ref)) < PSNR_TOL_DB
    r = nn.rate(enc.handle, sym, 512, 768)
    hist, ent, bpp, _ = O.rate(sym, 512, 768)
    assert np.array_equal(r.hist.astype(np.int64), hist)
    assert abs(float(r.bpp[0]) - float(O.rate(sym64, 512, 768)[2][0])) < BPP_TOL


@pytest.mark.parametrize("shape", [(1, 8, 8), (3, 8, 16), (2, 24, 40), (1, 136, 72), (5, 16, 8)])
def test_small_and_ragged_tiles(codec_factory, shape):
    """Smallest image, partial 16x8 tiles, batch sizes that are not multiples of anything."""
    n, h, w = shape
    img = synthetic_images(n, h, w, seed=h * 100 + w)
    eY, eC, dY, dC = make_weights("spread")
    for arith in ("tc_split", "simt_f32"):
        enc, dec = codec_factory("spread", arith)
        sym = enc(img)
        assert sym.shape == (n, h // 8, w // 8, 96)
        check_symbols(sym, O.encode(img, eY, eC, "f64"))
        rec = dec(sym)
        assert rec.shape == (n, h, w, 3)
        check_symbols(rec, O.decode(sym, dY, dC, "f64"))


@pytest.mark.parametrize("shape", [(2, 52, 44), (1, 9, 7), (3, 31, 33), (1, 255, 257), (1, 1, 1), (2, 100, 8), (1, 14, 130)])
def test_sizes_that_are_not_multiples_of_8(nn, codec_factory, shape):
    """TF SAME padding for odd sizes ((2,2) instead of (1,2)) and ceil at each stride-2 stage: both arithmetics against
    the fp64 oracle (mismatches only at rounding ties); the tensor-core kernels read the odd-sized activations through
    even-padded storage."""
    n, h, w = shape
    img = synthetic_images(n, h, w, seed=7)
    eY, eC, dY, dC = make_weights("spread")
    pre64 = O.encode_prequant(img, eY, eC, "f64")
    want = O.quantise(pre64)
    tie = np.abs(pre64 * 255.0 - np.floor(pre64 * 255.0) - 0.5)
    lam = SYMBOL_MISMATCH_LIMIT * want.size
    for arith in ("simt_f32", "tc_split"):
        enc, dec = codec_factory("spread", arith)
        sym, r = enc.encode_rate(img)
        assert sym.shape == want.shape == (n, -(-h // 8), -(-w // 8), 96)
        check_symbols(sym, want, tie, min_allow=int(lam + 4 * np.sqrt(lam) + 2))
        assert np.array_equal(r.hist.astype(np.int64), O.histogram(sym))
        assert enc.handle.arith == arith
        rec = dec(want)
        check_symbols(rec, O.decode(want, dY, dC, "f64"), min_allow=int(1e-4 * rec.size + 4 * np.sqrt(1e-4 * rec.size) + 2))
    # plane-level call (ProClass.run_model) on an odd size
    enc, _ = codec_factory("spread", "tc_split")
    planes = O.rgb_to_planes(img, "f32")
    got = enc.run_model(planes)
    for p in range(3):
        assert np.abs(got[p] - pre64[..., 32 * p:32 * p + 32]).max() < 2e-5


def test_micro_batches_and_device_api_are_equivalent(nn, codec_factory):
    import torch
    img = synthetic_images(7, 64, 48, seed=11)
    enc, dec = codec_factory("spread", "tc_split")
    ref_sym = enc(img)
    ref_rec = dec(ref_sym)
    for mb in (1, 2, 3):
        enc.handle.set_micro_batch(mb); dec.handle.set_micro_batch(mb)
        assert np.array_equal(enc(img), ref_sym)
        assert np.array_equal(dec(ref_sym), ref_rec)
    enc.handle.set_micro_batch(0); dec.handle.set_micro_batch(0)
    x = torch.from_numpy(img).cuda()
    sym_d = enc(x)
    assert sym_d.is_cuda and np.array_equal(sym_d.cpu().numpy(), ref_sym)
    rec_d = dec(sym_d)
    assert np.array_equal(rec_d.cpu().numpy(), ref_rec)
    r_d = nn.rate(enc.handle, sym_d, 64, 48)
    r_h = nn.rate(enc.handle, ref_sym, 64, 48)
    assert np.array_equal(r_d.hist.cpu().numpy().astype(np.uint32), r_h.hist)
    assert np.array_equal(r_d.bpp.cpu().numpy(), r_h.bpp)


@pytest.mark.parametrize("arith", ["tc_split", "simt_f32"])
def test_run_model_planes(codec_factory, arith):
    """ProClass.run_model (utils.py:19-24): float planes in, float planes out, weight sets (0,1,1)."""
    eY, eC, dY, dC = make_weights("spread")
    enc, dec = codec_factory("spread", arith)
    img = synthetic_images(2, 32, 48, seed=3)
    planes = O.rgb_to_planes(img, "f32")
    got = enc.run_model(planes)
    want = [O.base_encoder(planes[0], eY, "f64"), O.base_encoder(planes[1], eC, "f64"), O.base_encoder(planes[2], eC, "f64")]
    for g_, w_ in zip(got, want):
        assert g_.shape == w_.shape and g_.dtype == np.float32
        assert np.abs(g_ - w_).max() < 1e-5
    lat = [w_.astype(np.float32) for w_ in want]
    gotd = dec.run_model(lat)
    wantd = [O.base_decoder(lat[0], dY, "f64"), O.base_decoder(lat[1], dC, "f64"), O.base_decoder(lat[2], dC, "f64")]
    for g_, w_ in zip(gotd, wantd):
        assert g_.shape == w_.shape
        assert np.abs(g_ - w_).max() < 1e-5


def test_rate_edge_cases(nn, codec_factory):
    enc, _ = codec_factory("default", "tc_split")
    rng = np.random.default_rng(9)
    cases = {
        "zeros": np.zeros((2, 3, 5, 96), np.uint8),
        "all255": np.full((1, 4, 4, 96), 255, np.uint8),
        "uniform": rng.integers(0, 256, size=(3, 16, 24, 96), dtype=np.uint8),
        "peaked": np.minimum(rng.geometric(0.3, size=(2, 64, 96, 96)) - 1, 255).astype(np.uint8),
        "one_pixel": rng.integers(0, 256, size=(1, 1, 1, 96), dtype=np.uint8),
    }
    for name, lat in cases.items():
        n, lh, lw, _ = lat.shape
        r = nn.rate(enc.handle, lat, 8 * lh, 8 * lw)
        hist, ent, bpp, hg = O.rate(lat, 8 * lh, 8 * lw)
        assert np.array_equal(r.hist.astype(np.int64), hist), name
        assert np.array_equal(r.hist_global.astype(np.int64), hg), name
        assert int(r.hist.sum()) == lat.size
        assert np.abs(r.entropy_bits - ent).max() < 1e-5, name
        assert np.abs(r.bpp - bpp).max() < 1e-5, name
    # accumulation of the global histogram over micro-batches, then entropy of the reduced counts
    lat = cases["peaked"]
    acc = np.zeros((3, 256), np.uint64)
    nn.rate(enc.handle, lat[:1], hist_global=acc)
    nn.rate(enc.handle, lat[1:], hist_global=acc)
    want = O.histogram(lat).sum(axis=0)
    assert np.array_equal(acc.astype(np.int64), want)
    e = nn.entropy_from_counts(enc.handle, acc)
    assert np.abs(e - O.entropy_from_hist(want)).max() < 1e-5
    eg, bpp_g = nn.dist.global_rate(enc.handle, acc, 64, 96, 512, 768)
    assert abs(bpp_g - float(O.entropy_from_hist(want).sum() * 0.5)) < 1e-4


@pytest.mark.parametrize("arith,shape", [("tc_split", (5, 64, 96)), ("tc_split", (3, 200, 136)), ("tc_split", (2, 8, 8)),
                                         ("simt_f32", (2, 52, 44))])
def test_fused_encode_rate_equals_encode_then_rate(nn, codec_factory, arith, shape):
    """nnic_encode_rate counts the symbols inside the quantising kernel: latent, histograms, entropy and bpp must be
    bit-identical to encode followed by the standalone rate pass, for host and device buffers, with tiles that hang
    over the latent edge (200x136 -> 25x17 latent) and with micro-batches."""
    import torch
    enc, _ = codec_factory("spread", arith)
    n, hh, ww = shape
    x = synthetic_images(n, hh, ww, seed=21)
    lat = enc(x)
    want = nn.rate(enc.handle, lat, hh, ww)
    for mb in (0, 1, 2):
        enc.handle.set_micro_batch(mb)
        lat2, got = enc.encode_rate(x)
        assert np.array_equal(lat2, lat)
        assert np.array_equal(got.hist, want.hist)
        assert np.array_equal(got.hist_global, want.hist_global)
        assert np.array_equal(got.entropy_bits, want.entropy_bits)
        assert np.array_equal(got.bpp, want.bpp)
        assert int(got.hist.sum()) == lat.size
    enc.handle.set_micro_batch(0)
    acc = torch.zeros((3, 256), dtype=torch.int64, device="cuda")
    lat3, got3 = enc.encode_rate(torch.from_numpy(x).cuda(), hist_global=acc)
    lat3b, _ = enc.encode_rate(torch.from_numpy(x).cuda(), hist_global=acc)     # accumulates
    torch.cuda.synchronize()
    assert np.array_equal(lat3.cpu().numpy(), lat)
    assert np.array_equal(got3.hist.cpu().numpy().astype(np.uint32), want.hist)
    assert np.array_equal(acc.cpu().numpy().astype(np.uint64), 2 * want.hist_global)
    hist_o = O.histogram(lat)
    assert np.array_equal(want.hist.astype(np.int64), hist_o)


def test_argument_errors(nn, codec_factory):
    enc, dec = codec_factory("default", "tc_split")
    with pytest.raises(ValueError):
        enc(np.zeros((1, 8, 8, 4), np.uint8))
    with pytest.raises(ValueError):
        enc(np.zeros((1, 8, 8, 3), np.float32))
    with pytest.raises(ValueError):
        dec(np.zeros((1, 2, 2, 32), np.uint8))
    import torch
    xt = torch.zeros((2, 16, 24, 3), dtype=torch.uint8, device="cuda")
    for bad in (torch.empty((2, 2, 3, 32), dtype=torch.uint8, device="cuda"), torch.empty((2, 2, 3, 96), dtype=torch.float32, device="cuda"),
                torch.empty((2, 2, 3, 192), dtype=torch.uint8, device="cuda")[..., ::2]):
        with pytest.raises(ValueError):
            enc(xt, out=bad)
        with pytest.raises(ValueError):
            enc.encode_rate(xt, out=bad)
    with pytest.raises(ValueError):
        dec(torch.zeros((2, 2, 3, 96), dtype=torch.uint8, device="cuda"), out=torch.empty((2, 16, 24, 4), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        enc(torch.zeros((2, 16, 24, 3), dtype=torch.uint8))          # a CPU tensor is not silently copied
    # the C ABI rejects a host pointer passed as NNIC_MEM_DEVICE before anything is enqueued
    hx = np.zeros((1, 16, 24, 3), np.uint8); hl = np.zeros((1, 2, 3, 96), np.uint8)
    hnd = enc.handle
    assert hnd.lib.nnic_encode(hnd.h, hx.ctypes.data, 1, 16, 24, hl.ctypes.data, None, 1, None) == -1
    assert b"device memory" in hnd.lib.nnic_last_error(hnd.h) or b"CUDA pointer" in hnd.lib.nnic_last_error(hnd.h)
    assert hnd.lib.nnic_rate(hnd.h, hl.ctypes.data, 1, 2, 3, 16, 24, None, None, None, None, 1, None) == -1
    # replacing the weights of a handle that has already run takes effect on the next call
    img = synthetic_images(1, 32, 48, seed=13)
    sY, sC, _d2, _d3 = make_weights("spread")
    eY2, eC2, _d0, _d1 = make_weights("default")
    own = nn.Encoder(0)
    own.set_weights(0, sY); own.set_weights(1, sC)
    before = own(img)
    own.set_weights(0, eY2); own.set_weights(1, eC2)
    after = own(img)
    assert not np.array_equal(before, after) and np.array_equal(after, enc(img))
    own.set_weights(0, sY); own.set_weights(1, sC)
    assert np.array_equal(own(img), before)
    fresh = nn.Encoder(0)
    with pytest.raises(nn.NnicError, match="not set"):
        fresh(np.zeros((1, 8, 8, 3), np.uint8))
    with pytest.raises(nn.NnicError):
        nn.Handle(99)


# ---- BASELINE.json full-size configurations: size-independent properties + cross-arithmetic check ------
def _cross_check(nn, codec_factory, img, check_decode=True):
    enc_tc, dec_tc = codec_factory("spread", "tc_split")
    enc_ff, dec_ff = codec_factory("spread", "simt_f32")
    sym = enc_tc(img)
    sym_ff = enc_ff(img)
    check_symbols(sym, sym_ff)
    # batch independence: any image encoded alone gives the same bytes
    for i in (0, img.shape[0] - 1):
        assert np.array_equal(enc_tc(img[i:i + 1])[0], sym[i])
    # determinism
    assert np.array_equal(enc_tc(img), sym)
    r = nn.rate(enc_tc.handle, sym, img.shape[1], img.shape[2])
    assert int(r.hist.astype(np.int64).sum()) == sym.size
    assert np.array_equal(r.hist_global.astype(np.int64), r.hist.astype(np.int64).sum(axis=0))
    assert np.array_equal(r.hist[0, 1].astype(np.int64), np.bincount(sym[0, :, :, 32:64].ravel(), minlength=256))
    if check_decode:
        rec = dec_tc(sym)
        check_symbols(rec, dec_ff(sym))
        assert np.array_equal(dec_tc(sym[-1:])[0], rec[-1])
    return sym


def test_config2_kodak_batch(nn, codec_factory):
    """24 x 768x512 full encode + rate + decode (config 2 shape)."""
    _cross_check(nn, codec_factory, synthetic_images(24, 512, 768, seed=2))


def config2_batch():
    """SURVEY.md 8d, C2 recipe (tests/parity_fixture.py::c2_images)."""
    import parity_fixture as PF
    return PF.c2_images()


def test_config2_recipe_against_oracle(nn, codec_factory):
    """BASELINE.json config 2 at full size (24 x 768x512) against the oracle itself: symbols (ties only), histograms,
    bpp and reconstruction PSNR, through the fused encode + rate call and the decoder."""
    img = config2_batch()
    assert img.shape == (24, 512, 768, 3)
    eY, eC, dY, dC = make_weights("spread")
    enc, dec = codec_factory("spread", "tc_split")
    sym, r = enc.encode_rate(img)
    pre64 = O.encode_prequant(img, eY, eC, "f64")
    sym64 = O.quantise(pre64)
    tie = np.abs(pre64 * 255.0 - np.floor(pre64 * 255.0) - 0.5)
    check_symbols(sym, sym64, tie)
    hist, ent, bpp, hg = O.rate(sym, 512, 768)
    assert np.array_equal(r.hist.astype(np.int64), hist)
    assert np.array_equal(r.hist_global.astype(np.int64), hg)
    assert np.abs(r.bpp - O.rate(sym64, 512, 768)[2]).max() < BPP_TOL
    rec = dec(sym64)
    rec_ref = O.decode(sym64, dY, dC, "f32")
    check_symbols(rec, rec_ref)
    for i in range(24):
        assert abs(O.psnr(img[i], rec[i]) - O.psnr(img[i], rec_ref[i])) < PSNR_TOL_DB


# Configs 3, 4 and 5 at their full sizes (and config 2 with both weight sets) are compared with the oracle in
# tests/test_gpu_fullsize.py; the cross-arithmetic check below stays as a cheap GPU-vs-GPU guard on a config-5-shaped shard.
def test_config5_cross_arithmetic(nn, codec_factory):
    sym = _cross_check(nn, codec_factory, synthetic_images(64, 256, 256, seed=5), check_decode=False)
    assert sym.shape == (64, 32, 32, 96)


def test_compress_uncompress_directories(nn, codec_factory, tmp_path):
    """Encoder.compress / Decoder.uncompress (encoder.py:49-51, decoder.py:50-52, utils.py:30-62): a directory of
    images -> `<dir>_compressed/*.png` (packed latents) -> `<dir>_uncompressed/*.png`, equal to the in-memory path
    and to the oracle's packing; weights go through save()/load() as `path+'Y'`, `path+'CbCr'`."""
    from PIL import Image
    enc, dec = codec_factory("spread", "tc_split")
    imgs = synthetic_images(6, 64, 96, seed=31)           # 6 images: one full batch of 4 and a rest of 2
    src = tmp_path / "kodak"
    src.mkdir()
    for i, im in enumerate(imgs):
        nn.save_img(im, str(src), f"im{i:02d}")
    enc.save(str(tmp_path / "encoder")); dec.save(str(tmp_path / "decoder"))
    enc2, dec2 = nn.Encoder(0), nn.Decoder(0)
    cdir = enc2.compress(str(src), str(tmp_path / "encoder"))
    assert cdir == str(src) + "_compressed"
    udir = dec2.uncompress(cdir, str(tmp_path / "decoder"))
    assert udir == str(src) + "_uncompressed"
    lat = enc(imgs)
    rec = dec(lat)
    packed = O.pack_latent(lat)
    for i in range(6):
        c = np.array(Image.open(f"{cdir}/im{i:02d}.png"))
        u = np.array(Image.open(f"{udir}/im{i:02d}.png"))
        assert c.shape == (4 * 8, 8 * 12, 3) and np.array_equal(c, packed[i])
        assert u.shape == (64, 96, 3) and np.array_equal(u, rec[i])
    # a directory with two image sizes is processed one image at a time
    nn.save_img(synthetic_images(1, 32, 48, seed=32)[0], str(src), "small")
    enc2.compress(str(src))
    small = np.array(Image.open(f"{cdir}/small.png"))
    assert np.array_equal(nn.unpack_latent(small[None]), enc(synthetic_images(1, 32, 48, seed=32)))


@pytest.mark.parametrize("shape", [(3, 72, 40), (1, 8, 8), (2, 136, 264), (2, 45, 67)])
def test_outputs_stay_inside_their_buffers(nn, codec_factory, shape):
    """Every output is placed in the middle of a larger sentinel-filled device buffer: the kernels (partial tiles at
    the right / bottom edge, 256-bit stores, the histogram flush) must not write one byte outside it."""
    import torch
    enc, dec = codec_factory("spread", "tc_split")
    n, hh, ww = shape
    lh, lw = -(-hh // 8), -(-ww // 8)
    x = torch.from_numpy(synthetic_images(n, hh, ww, seed=41)).cuda()
    guard = 4096

    def framed(numel, dtype, fill):
        big = torch.full((numel + 2 * guard,), fill, dtype=dtype, device="cuda")
        return big, big[guard:guard + numel]

    big_lat, lat = framed(n * lh * lw * 96, torch.uint8, 0xA5)
    big_pre, pre = framed(n * lh * lw * 96, torch.float32, -7.0)
    big_rgb, rgb = framed(n * 8 * lh * 8 * lw * 3, torch.uint8, 0x5A)
    big_hg, hg = framed(3 * 256, torch.int64, -1)
    lat = lat.view(n, lh, lw, 96); rgb = rgb.view(n, 8 * lh, 8 * lw, 3)
    hg.zero_()
    hg = hg.view(3, 256)
    # nnic_encode with the optional pre-quantisation output, then the fused encode + rate, then decode
    h = enc.handle
    h.check(h.lib.nnic_encode(h.h, x.data_ptr(), n, hh, ww, lat.data_ptr(), pre.data_ptr(), 1, None), "nnic_encode")
    lat_a = lat.clone()
    lat2, r = enc.encode_rate(x, out=lat, hist_global=hg)
    dec(lat, out=rgb)
    torch.cuda.synchronize()
    for big, fill in ((big_lat, 0xA5), (big_pre, -7.0), (big_rgb, 0x5A), (big_hg, -1)):
        assert bool((big[:guard] == fill).all()) and bool((big[-guard:] == fill).all())
    assert torch.equal(lat_a, lat2) and int(hg.sum()) == lat.numel()
    want_lat = enc(x.cpu().numpy())
    assert np.array_equal(lat.cpu().numpy(), want_lat)
    assert np.array_equal(rgb.cpu().numpy(), dec(want_lat))
    assert np.array_equal(np.round(pre.view(n, lh, lw, 96).cpu().numpy() * np.float32(255)).astype(np.uint8), want_lat)


@pytest.mark.parametrize("wname", ["default", "spread"])
def test_fp16_decoder_meets_the_reconstruction_tolerance(nn, wname):
    """Optional decoder arithmetic NNIC_DECODE_FP16 (one fp16 product per MAC): not byte-identical, but inside
    BASELINE.json's reconstruction criterion -- PSNR against the source within 0.01 dB of the reference decoder's, every
    differing byte +-1 -- on kodim21 (config 1), the golden crops and a ragged-tile size; the exact decoder on the same
    handle type is unaffected, and the encoder ignores the setting."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    eY, eC, dY, dC = make_weights(wname)
    enc = nn.Encoder(0)
    enc.set_weights(0, eY); enc.set_weights(1, eC)
    dec_exact, dec_fast = nn.Decoder(0), nn.Decoder(0, precision="fp16")
    assert dec_fast.handle.decode_precision == "fp16" and dec_exact.handle.decode_precision == "split"
    for d in (dec_exact, dec_fast):
        d.set_weights(0, dY); d.set_weights(1, dC)
    kodim = np.array(Image.open(os.path.join(GOLDEN, "kodim21.png")))[None]
    for img in (kodim, load_golden("imagenet_patches")["input"], synthetic_images(3, 72, 40, seed=51)):
        sym = enc(img)
        rec_ref = O.decode(sym, dY, dC, "f32")
        rec_exact, rec_fast = dec_exact(sym), dec_fast(sym)
        check_symbols(rec_exact, rec_ref)
        d = rec_fast.astype(int) - rec_ref.astype(int)
        assert np.abs(d).max() <= 1
        assert (d != 0).mean() < 0.12
        for i in range(img.shape[0]):
            assert abs(O.psnr(img[i], rec_fast[i]) - O.psnr(img[i], rec_ref[i])) < PSNR_TOL_DB
            assert O.psnr(rec_ref[i], rec_fast[i]) > 55.0
    # the encoder of a handle set to fp16 decoding still produces the exact symbols
    enc2 = nn.Encoder(0)
    enc2.handle.set_decode_precision("fp16")
    enc2.set_weights(0, eY); enc2.set_weights(1, eC)
    assert np.array_equal(enc2(kodim), enc(kodim))
    with pytest.raises(nn.NnicError):
        dec_fast.handle.set_decode_precision(7)


def test_load_reads_tensorflow_checkpoints(nn, codec_factory, tmp_path):
    """ProClass.load(path) with `<path>Y.index` / `<path>CbCr.index` present (what the reference's save_weights leaves
    there) gives the same codec as installing the arrays directly."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from tf_bundle_writer import write_bundle
    eY, eC, dY, dC = make_weights("spread")
    for prefix, w in (("encoderY", eY), ("encoderCbCr", eC), ("decoderY", dY), ("decoderCbCr", dC)):
        write_bundle(str(tmp_path / prefix), {f"{k}/.ATTRIBUTES/VARIABLE_VALUE": v for k, v in w.items()})
    enc = nn.Encoder(0).load(str(tmp_path / "encoder"))
    dec = nn.Decoder(0).load(str(tmp_path / "decoder"))
    enc_ref, dec_ref = codec_factory("spread", "tc_split")
    img = synthetic_images(2, 32, 48, seed=61)
    sym = enc(img)
    assert np.array_equal(sym, enc_ref(img))
    assert np.array_equal(dec(sym), dec_ref(sym))


def test_random_shapes_against_oracle(nn, codec_factory):
    """Seeded random batch / image sizes (1-7 images; half of them multiples of 8 up to 296 x 344, half arbitrary sizes up
    to 300 x 340: partial tiles, odd sizes at every stride-2 stage) against the fp64 oracle: every mismatching symbol must be a rounding tie (+-1, inside the tie band), and over all
    shapes together the mismatch fraction must meet the 1e-4 criterion (single small images get Poisson slack).  The
    fused histogram must equal the histogram of the symbols; the decoder is compared with the FFMA arithmetic."""
    eY, eC, dY, dC = make_weights("spread")
    enc_tc, dec_tc = codec_factory("spread", "tc_split")
    _e, dec_ff = codec_factory("spread", "simt_f32")
    rng = np.random.default_rng(77)
    bad = total = 0
    for k in range(20):
        n, h, w = int(rng.integers(1, 8)), 8 * int(rng.integers(1, 38)), 8 * int(rng.integers(1, 44))
        if k % 2:
            h, w = int(rng.integers(1, 301)), int(rng.integers(1, 341))
        img = synthetic_images(n, h, w, seed=int(rng.integers(0, 1 << 30)))
        sym, r = enc_tc.encode_rate(img)
        pre64 = O.encode_prequant(img, eY, eC, "f64")
        sym64 = O.quantise(pre64)
        tie = np.abs(pre64 * 255.0 - np.floor(pre64 * 255.0) - 0.5)
        lam = SYMBOL_MISMATCH_LIMIT * sym.size
        bad += check_symbols(sym, sym64, tie, min_allow=int(lam + 4 * np.sqrt(lam) + 2))
        total += sym.size
        assert np.array_equal(r.hist.astype(np.int64), O.histogram(sym)), (n, h, w)
        rec, rec_ff = dec_tc(sym64), dec_ff(sym64)
        assert np.abs(rec.astype(int) - rec_ff.astype(int)).max() <= 1 and (rec != rec_ff).mean() < 3e-4, (n, h, w)
    assert bad <= SYMBOL_MISMATCH_LIMIT * total, f"{bad} of {total} symbols differ from the fp64 oracle"


def test_cluster_multicast_variant_is_bit_identical(nn, monkeypatch):
    """NNIC_TC_CLUSTER=2 (CTA pairs sharing every weight tile through TMA multicast, opt-in: measured no faster,
    profiles/r1_cluster_multicast_ab.log) must give the same bytes as the default kernels, including an odd number of
    work items (the last cluster runs with one idle CTA) and a pair that straddles the Y / CbCr weight sets."""
    eY, eC, dY, dC = make_weights("spread")

    def codec():
        e, d = nn.Encoder(0), nn.Decoder(0)
        e.set_weights(0, eY); e.set_weights(1, eC); d.set_weights(0, dY); d.set_weights(1, dC)
        return e, d
    enc0, dec0 = codec()
    monkeypatch.setenv("NNIC_TC_CLUSTER", "2")
    enc1, dec1 = codec()
    for shape in ((1, 8, 8), (1, 72, 40), (3, 136, 264), (5, 64, 96), (2, 45, 67)):
        img = synthetic_images(*shape, seed=71)
        sym0, r0 = enc0.encode_rate(img)
        sym1, r1 = enc1.encode_rate(img)
        assert np.array_equal(sym0, sym1) and np.array_equal(r0.hist, r1.hist), shape
        assert np.array_equal(dec0(sym0), dec1(sym0)), shape


def test_fused_dconv7_dconv8_is_bit_identical(nn, monkeypatch):
    """The decoder's default path never writes dconv7's output: the layer's epilogue feeds dconv8's tap-response GEMM on chip
    (tc_conv_patch.cu FUSE8) and k_dconv8_gather sums the responses (decoder.py:16-17,30-32,45-48).  It must give the bytes,
    the pre-quantisation floats and the plane outputs of the unfused kernels (NNIC_FUSE_D78=0: dconv7 -> memory ->
    k_tc_dconv8), on ragged tiles, one-pixel latents, several images and both weight sets."""
    import torch
    _, _, dY, dC = make_weights("spread")

    def decoder():
        d = nn.Decoder(0)
        d.set_weights(0, dY); d.set_weights(1, dC)
        return d
    dec1 = decoder()
    monkeypatch.setenv("NNIC_FUSE_D78", "0")
    dec0 = decoder()
    rng = np.random.default_rng(5)
    for shape in ((1, 1, 1), (1, 9, 5), (3, 17, 33), (5, 8, 12), (2, 6, 9), (2, 64, 96)):
        n, lh, lw = shape
        sym = rng.integers(0, 256, size=(n, lh, lw, 96)).astype(np.uint8)
        sym[rng.random(sym.shape) < 0.5] = 0
        r0, p0 = dec0(sym, return_prequant=True)
        r1, p1 = dec1(sym, return_prequant=True)
        assert np.array_equal(r0, r1), shape
        assert np.array_equal(p0, p1), shape
        lat = [rng.random((n, lh, lw, 32)).astype(np.float32) for _ in range(3)]
        for a, b in zip(dec0.run_model(lat), dec1.run_model(lat)):
            assert np.array_equal(a, b), shape
        # device buffers, output at an odd byte offset
        xs = torch.from_numpy(sym).cuda()
        buf = torch.zeros(n * 8 * lh * 8 * lw * 3 + 8, dtype=torch.uint8, device="cuda")
        out = buf[1:1 + n * 8 * lh * 8 * lw * 3].view(n, 8 * lh, 8 * lw, 3)
        dec1(xs, out=out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), r0), shape
        assert int(buf[0]) == 0 and int(buf[-7:].sum()) == 0


def test_pinned_weight_tiles_are_bit_identical(nn, monkeypatch):
    """The nine-tap layers (conv3, conv4, dconv5, dconv6) keep seven weight tiles in shared memory for all items of a weight
    set and stream two (tc_conv_patch.cu PIN); NNIC_TC_PIN=0 streams every tile per item.  Same bytes, including CTAs whose
    item sequence crosses the Y -> CbCr weight-set boundary and CTAs with a single item."""
    eY, eC, dY, dC = make_weights("spread")

    def codec():
        e, d = nn.Encoder(0), nn.Decoder(0)
        e.set_weights(0, eY); e.set_weights(1, eC); d.set_weights(0, dY); d.set_weights(1, dC)
        return e, d
    enc1, dec1 = codec()                       # default: the residual layers conv4 / dconv6
    monkeypatch.setenv("NNIC_TC_PIN", "2")     # all four nine-tap layers
    enc2, dec2 = codec()
    monkeypatch.setenv("NNIC_TC_PIN", "0")
    enc0, dec0 = codec()
    for shape in ((1, 8, 8), (1, 72, 40), (3, 136, 264), (5, 256, 384), (2, 45, 67), (7, 128, 128)):
        img = synthetic_images(*shape, seed=73)
        sym0, r0 = enc0.encode_rate(img)
        for e, d in ((enc1, dec1), (enc2, dec2)):
            sym1, r1 = e.encode_rate(img)
            assert np.array_equal(sym0, sym1) and np.array_equal(r0.hist, r1.hist), shape
            assert np.array_equal(dec0(sym0), d(sym0)), shape


def test_back_to_back_large_decodes_are_stable(nn):
    """Device-buffer decodes of 1080p-class latents issued back to back without host synchronisation, alternating two inputs:
    the pattern of bench.py's config-4 loop, which the packed-fp32 epilogues of round 2 failed intermittently on some boxes
     while every other test passed.  Each input must give the same bytes every time."""
    import torch
    dec = nn.Decoder(0)
    dec.init_random()
    g = torch.Generator(device="cuda").manual_seed(3)
    lats = []
    for _ in range(2):
        u = torch.rand((8, 135, 240, 96), device="cuda", generator=g)
        lats.append((torch.log1p(-u) / -0.08).clamp_(0, 255).to(torch.uint8))
    outs = [torch.empty((8, 1080, 1920, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    refs = []
    for lat in lats:
        refs.append(dec(lat).clone())
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(8):
            dec(lats[i & 1], out=outs[i & 1])
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(outs[k], refs[k]), (rep, k)


def test_bench_line_has_the_contract_keys():
    """`python bench.py` on one GPU: one JSON line with value, e2e (host copies counted), roofline, cpu_baseline, clocks and a
    positive launch count; the roofline kernel's share comes from the per-launch event timing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "5", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "encode+decode megapixels/sec" and d["unit"] == "MP/s" and d["n_gpus"] == 1 and d["value"] > 100
    assert d["steps"] == 5 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"]["value"] > 100 and d["e2e"]["h2d_bytes_per_step"] >= 24 * 512 * 768 * 3
    assert d["e2e"]["d2h_bytes_per_step"] >= 24 * 512 * 768 * 3 + 24 * 64 * 96 * 96
    assert d["gpu_launches"] >= 5 * 10
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    assert rf["kernel"] in d["kernels"] and rf["traffic"] is not None
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert "sm_mhz" in d["clocks"] and "workload" in d["config"] and "model" not in d["config"]


def test_device_calls_can_be_captured_in_a_cuda_graph(nn, codec_factory):
    """With device buffers the library only enqueues kernels and memsets on the caller's stream, so once the scratch
    buffers have their size (one warm-up call) encode_rate + decode can be captured in a CUDA graph and replayed on new
    input data with identical results (tools/graph_latency.py: 237 -> 213 us for one 768x512 image)."""
    import torch
    enc, dec = codec_factory("spread", "tc_split")
    x = torch.from_numpy(synthetic_images(2, 64, 96, seed=81)).cuda()
    x2 = torch.from_numpy(synthetic_images(2, 64, 96, seed=82)).cuda()
    lat = torch.empty((2, 8, 12, 96), dtype=torch.uint8, device="cuda")
    rgb = torch.empty((2, 64, 96, 3), dtype=torch.uint8, device="cuda")
    hg = torch.zeros((3, 256), dtype=torch.int64, device="cuda")
    xin = x.clone()

    def step():
        hg.zero_()
        _, r = enc.encode_rate(xin, out=lat, hist_global=hg)
        dec(lat, out=rgb)
        return r
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        r = step()
    for inp in (x2, x):
        xin.copy_(inp)
        lat.zero_(); rgb.zero_()
        graph.replay()
        torch.cuda.synchronize()
        want_lat = enc(inp.cpu().numpy())
        assert np.array_equal(lat.cpu().numpy(), want_lat)
        assert np.array_equal(rgb.cpu().numpy(), dec(want_lat))
        assert np.array_equal(r.hist.cpu().numpy().astype(np.int64), O.histogram(want_lat))
        assert np.array_equal(hg.cpu().numpy(), O.histogram(want_lat).sum(axis=0))


def test_c_program_on_the_abi_matches_the_python_layer(nn, tmp_path):
    """examples/c_abi_demo.c drives libnnic.so from plain C with host buffers (nnic_create, nnic_set_weights,
    nnic_encode_rate, nnic_decode).  The test rebuilds its LCG weights and image in Python and expects the same latent and
    reconstruction bytes (FNV-1a), bpp and symbol count."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "neural_network_image_compression_b200")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-std=c99", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_demo.c"),
                    "-L", lib_dir, "-lnnic", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = dict(re.findall(r"\b(latent_fnv|recon_fnv|default_latent_fnv|default_recon_fnv) ([0-9a-f]{16})", out.stdout))
    assert "channel_rows_ok 1" in out.stdout
    bpp_c = [float(v) for v in re.search(r"bpp (\S+) (\S+)", out.stdout).groups()]

    state = 12345

    def draws(n):                                   # the program's LCG: float32((state >> 8) * 2^-24)
        nonlocal state
        vals = np.empty(n, np.float32)
        for i in range(n):
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
            vals[i] = np.float32(state >> 8) * np.float32(1.0 / 16777216.0)
        return vals
    enc_l = ((5, 1, 32), (5, 32, 64), (3, 64, 64), (3, 64, 64), (5, 64, 32))
    dec_l = ((5, 32, 64), (3, 64, 64), (3, 64, 64), (5, 64, 64), (5, 64, 1))
    sets = []
    for s in range(4):
        w = {}
        names = [layer[0] for layer in nn.weights.layers_of("encoder" if s < 2 else "decoder")]
        for name, (k, cin, cout) in zip(names, enc_l if s < 2 else dec_l):
            limit = np.float32(1.6) * np.sqrt(np.float32(6.0) / np.float32((cin + cout) * k * k)).astype(np.float32)
            kern = (np.float32(2.0) * draws(k * k * cin * cout) - np.float32(1.0)) * limit
            w[name + "/kernel"] = kern.reshape((k, k, cin, cout) if s < 2 else (k, k, cout, cin)).astype(np.float32)
            w[name + "/bias"] = ((np.float32(2.0) * draws(cout) - np.float32(1.0)) * np.float32(0.05)).astype(np.float32)
        sets.append(w)
    n_, h_, w_ = 2, 64, 96
    img = np.empty((n_, h_, w_, 3), np.uint8)
    for n in range(n_):
        for y in range(h_):
            for x in range(w_):
                for c in range(3):
                    state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
                    u = float(np.float32(state >> 8) * np.float32(1.0 / 16777216.0))
                    img[n, y, x, c] = int(128.0 + 100.0 * np.sin(0.11 * x + 0.07 * y * (c + 1) + n) + 20.0 * u) & 0xFF
    enc, dec = nn.Encoder(0), nn.Decoder(0)
    enc.set_weights(0, sets[0]); enc.set_weights(1, sets[1]); dec.set_weights(0, sets[2]); dec.set_weights(1, sets[3])
    lat, r = enc.encode_rate(img)
    rec = dec(lat)

    def fnv(a):
        hsh = 1469598103934665603
        for b in a.tobytes():
            hsh = ((hsh ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return f"{hsh:016x}"
    assert got["latent_fnv"] == fnv(lat) and got["recon_fnv"] == fnv(rec)
    assert np.allclose(bpp_c, r.bpp, atol=1e-6)
    # second pass of the program: nnic_init_random(h, set, 11 + set) == the "default" weight sets of the parity tests
    eY, eC, dY, dC = make_weights("default")
    enc.set_weights(0, eY); enc.set_weights(1, eC); dec.set_weights(0, dY); dec.set_weights(1, dC)
    lat_d = enc(img)
    assert got["default_latent_fnv"] == fnv(lat_d) and got["default_recon_fnv"] == fnv(dec(lat_d))


def test_init_random_from_the_library(nn, codec_factory):
    """Encoder().init_random() / Decoder().init_random() draw the networks inside libnnic (nnic_init_random_scaled): same
    bytes out as installing weights.glorot_uniform(kind, seed) through set_weights, for the default and a scaled draw."""
    img = synthetic_images(2, 40, 56, seed=91)
    for wname, gain, br in (("default", 1.0, 0.0), ("spread", 1.6, 0.05)):
        enc_ref, dec_ref = codec_factory(wname, "tc_split")
        enc, dec = nn.Encoder(0).init_random(gain=gain, bias_range=br), nn.Decoder(0).init_random(gain=gain, bias_range=br)
        sym = enc(img)
        assert np.array_equal(sym, enc_ref(img)) and np.array_equal(dec(sym), dec_ref(sym))
        eY = make_weights(wname)[0]
        assert all(np.array_equal(enc.weights[0][k], eY[k]) for k in eY)


def test_rate_channels_and_large_counts(nn, codec_factory):
    """Per-feature-channel table [96,256]: equal to a NumPy bincount per channel, its 32-row groups add up to the plane
    histogram of nnic_rate, it accumulates, host and device buffers agree.  Entropy of counts beyond 2^24 (the all-rank
    totals of config 5 are ~2e9 per plane) follows p = count / total with the exact integer total."""
    import torch
    enc, _ = codec_factory("default", "tc_split")
    rng = np.random.default_rng(19)
    lat = np.minimum(rng.geometric(0.2, size=(5, 33, 17, 96)) - 1, 255).astype(np.uint8)
    lat[:, :, :, 7] = 0; lat[:, :, :, 40] = 255
    want = np.stack([np.bincount(lat[..., c].ravel(), minlength=256) for c in range(96)]).astype(np.int64)
    got = nn.rate_channels(enc.handle, lat)
    assert got.shape == (96, 256) and np.array_equal(got.astype(np.int64), want)
    r = nn.rate(enc.handle, lat)
    assert np.array_equal(got.astype(np.int64).reshape(3, 32, 256).sum(axis=1), r.hist_global.astype(np.int64))
    acc = torch.zeros((96, 256), dtype=torch.int64, device="cuda")
    nn.rate_channels(enc.handle, torch.from_numpy(lat[:2]).cuda(), hist_channels=acc)
    nn.rate_channels(enc.handle, torch.from_numpy(lat[2:]).cuda(), hist_channels=acc)
    assert np.array_equal(acc.cpu().numpy(), want)
    tiny = rng.integers(0, 256, size=(1, 1, 1, 96), dtype=np.uint8)
    assert int(nn.rate_channels(enc.handle, tiny).sum()) == 96
    # counts far beyond 2^24
    big = np.zeros((3, 256), np.uint64)
    big[0, :4] = [2_000_000_001, 1_000_000_007, 16_777_217, 3]
    big[1] = rng.integers(1 << 24, 1 << 33, size=256).astype(np.uint64)
    big[2, 200] = (1 << 40) + 12345
    e = nn.entropy_from_counts(enc.handle, big)
    p = big.astype(np.float64) / big.sum(axis=1, keepdims=True)
    want_e = (p * -np.log2(np.clip(p, 1e-5, 1.0))).sum(axis=1)
    assert np.abs(e - want_e).max() < 2e-6
    e_dev = nn.entropy_from_counts(enc.handle, torch.from_numpy(big.astype(np.int64)).cuda())
    assert np.array_equal(e_dev.cpu().numpy(), e)


def test_rate_buffers_are_validated(nn, codec_factory):
    """the C side reads and writes 768 (or 24 576) 64-bit counters at hist_global / hist_channels; every other
    dtype, shape, layout or memory side is rejected in the Python layer, and CPU torch tensors take the host path."""
    import torch
    enc, _ = codec_factory("default", "tc_split")
    lat = np.zeros((2, 3, 5, 96), np.uint8)
    lat_d = torch.from_numpy(lat).cuda()
    x = synthetic_images(2, 24, 40, seed=5)
    bad_host = (np.zeros((3, 256), np.uint32), np.zeros((3, 256), np.int32), np.zeros((3, 256), np.float64), np.zeros((256, 3), np.uint64),
                np.zeros((3, 512), np.uint64)[:, ::2], np.zeros((768,), np.uint64), torch.zeros((3, 256), dtype=torch.int64, device="cuda"))
    for bad in bad_host:
        with pytest.raises((ValueError, TypeError)):
            nn.rate(enc.handle, lat, hist_global=bad)
        with pytest.raises((ValueError, TypeError)):
            enc.encode_rate(x, hist_global=bad)
    bad_dev = (torch.zeros((3, 256), dtype=torch.int32, device="cuda"), torch.zeros((3, 256), dtype=torch.float32, device="cuda"),
               torch.zeros((3, 512), dtype=torch.int64, device="cuda")[:, ::2], np.zeros((3, 256), np.uint64),
               torch.zeros((3, 256), dtype=torch.int64))
    for bad in bad_dev:
        with pytest.raises((ValueError, TypeError)):
            nn.rate(enc.handle, lat_d, hist_global=bad)
        with pytest.raises((ValueError, TypeError)):
            enc.encode_rate(torch.from_numpy(x).cuda(), hist_global=bad)
    with pytest.raises(ValueError):
        nn.rate(enc.handle, lat.astype(np.int8))
    with pytest.raises(ValueError):
        nn.rate(enc.handle, lat_d.to(torch.int8))
    with pytest.raises(ValueError):
        nn.rate_channels(enc.handle, lat, hist_channels=np.zeros((3, 256), np.uint64))
    with pytest.raises(ValueError):
        nn.entropy_from_counts(enc.handle, np.zeros((3, 256), np.float32))
    with pytest.raises(ValueError):
        nn.entropy_from_counts(enc.handle, np.zeros((3, 128), np.uint64))
    # CPU torch tensors are host memory: they take the host path and are written in place
    hg = torch.zeros((3, 256), dtype=torch.int64)
    r = nn.rate(enc.handle, torch.from_numpy(lat), hist_global=hg)
    assert int(hg.sum()) == lat.size and int(hg[0, 0]) == lat.size // 3 and isinstance(r.hist, np.ndarray)
    e_host = nn.entropy_from_counts(enc.handle, hg)
    eg, _bpp = nn.dist.global_rate(enc.handle, hg, 3, 5, 24, 40)
    assert np.array_equal(np.asarray(e_host), eg)
    # micro-batch override beyond the plane limit of the kernels' grids is clamped, not passed through
    want = enc(x)
    enc.handle.set_micro_batch(1 << 30)
    assert np.array_equal(enc(x), want)
    enc.handle.set_micro_batch(0)


def test_graph_codec_matches_direct_calls(nn, codec_factory):
    """GraphCodec (CUDA-graph replay of encode_rate + decode for a fixed shape) returns the bytes of the direct calls for
    device, pinned-host and NumPy inputs; and a steady-state direct call encodes no tensor map (they are cached)."""
    import torch
    enc, dec = codec_factory("spread", "tc_split")
    gc = nn.GraphCodec(enc, dec, 2, 72, 104)
    for seed in (101, 102, 103):
        img = synthetic_images(2, 72, 104, seed=seed)
        want_lat = enc(img)
        want_rec = dec(want_lat)
        src = (torch.from_numpy(img).cuda(), torch.from_numpy(img).pin_memory(), img)[seed - 101]
        lat, r, rgb = gc.run(src)
        torch.cuda.synchronize()
        assert np.array_equal(lat.cpu().numpy(), want_lat) and np.array_equal(rgb.cpu().numpy(), want_rec)
        assert np.array_equal(r.hist.cpu().numpy().astype(np.int64), O.histogram(want_lat))
        assert np.array_equal(gc.hist_global.cpu().numpy(), O.histogram(want_lat).sum(axis=0))
    x = torch.from_numpy(synthetic_images(2, 72, 104, seed=104)).cuda()
    lat = torch.empty((2, 9, 13, 96), dtype=torch.uint8, device="cuda")
    rgb = torch.empty((2, 72, 104, 3), dtype=torch.uint8, device="cuda")
    enc(x, out=lat); dec(lat, out=rgb)
    before = enc.handle.lib.nnic_tensor_map_encodes(enc.handle.h) + dec.handle.lib.nnic_tensor_map_encodes(dec.handle.h)
    for _ in range(3):
        enc(x, out=lat); dec(lat, out=rgb)
    after = enc.handle.lib.nnic_tensor_map_encodes(enc.handle.h) + dec.handle.lib.nnic_tensor_map_encodes(dec.handle.h)
    assert after == before


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 9, 7), (3, 40, 56), (1, 64, 96)])
def test_encoder_input_at_any_byte_alignment(nn, codec_factory, shape):
    """conv1 fetches the RGB patch with aligned 32-bit loads and extracts the bytes in shared memory: an input tensor that
    starts at any byte offset (a slice of a larger uint8 buffer) and ends flush with its allocation must give the bytes of
    the aligned tensor -- the words that straddle the ends of the buffer are assembled from in-range byte loads."""
    import torch
    enc, _ = codec_factory("spread", "tc_split")
    n, hh, ww = shape
    img = synthetic_images(n, hh, ww, seed=hh + 3)
    want = enc(img)
    count = img.size
    for off in (1, 2, 3, 5):
        flat = torch.zeros(count + off, dtype=torch.uint8, device="cuda")      # the view ends exactly at the allocation's last byte
        flat[off:] = torch.from_numpy(img.reshape(-1)).cuda()
        x = flat[off:].view(n, hh, ww, 3)
        assert x.data_ptr() % 4 == off % 4
        assert np.array_equal(enc(x).cpu().numpy(), want), (shape, off)


def test_saturation_of_the_split_representation_is_detectable(nn, codec_factory):
    """activations are stored as fp16 pairs of v*16 and saturate at |v| > 4094, silently.  With the test weight sets
    nothing saturates; with one kernel blown up by 1e5 the debug counter reports it (and the fp32 FFMA arithmetic differs)."""
    img = synthetic_images(2, 64, 96, seed=17)
    enc, dec = codec_factory("spread", "tc_split")
    lat = enc(img); dec(lat)
    assert enc.handle.saturated_activations() == 0 and dec.handle.saturated_activations() == 0
    eY, eC, _dY, _dC = make_weights("spread")
    big = nn.Encoder(0)
    for i, w in enumerate((eY, eC)):
        w2 = {k: (v * np.float32(1.0e5) if k.startswith("conv2/kernel") else v) for k, v in w.items()}
        big.set_weights(i, w2)
    big(img)
    assert big.handle.saturated_activations() > 0


def test_heavy_tailed_weights_against_oracle(nn):
    """Trained networks do not look like a glorot draw: per-channel gains that span two orders of magnitude and biases of
    +-0.5 exercise the power-of-two weight scaling (max |w| just below 32768) and the hi/lo split of small weights next to
    large ones.  Symbols and reconstruction bytes against the fp64 oracle, ties only."""
    from neural_network_image_compression_b200 import weights as Wt
    rng = np.random.default_rng(123)

    def heavy(kind, seed):
        w = Wt.glorot_uniform(kind, seed, 1.0, 0.0)
        for name, _k, _s, _cin, cout in Wt.layers_of(kind):
            gains = np.exp(rng.uniform(np.log(0.15), np.log(6.0), size=cout)).astype(np.float32)
            gains /= np.float32(np.sqrt(np.mean(gains ** 2)))                 # keep the layer's overall scale
            kern = w[name + "/kernel"]
            w[name + "/kernel"] = (kern * gains if kind == "encoder" else kern * gains[None, None, :, None]).astype(np.float32)
            w[name + "/bias"] = rng.uniform(-0.5, 0.5, size=cout).astype(np.float32) * np.float32(0.2)
        return w
    eY, eC, dY, dC = heavy("encoder", 31), heavy("encoder", 32), heavy("decoder", 33), heavy("decoder", 34)
    enc, dec = nn.Encoder(0), nn.Decoder(0)
    enc.set_weights(0, eY); enc.set_weights(1, eC); dec.set_weights(0, dY); dec.set_weights(1, dC)
    img = synthetic_images(3, 128, 192, seed=55)
    sym, pre = enc(img, return_prequant=True)
    pre64 = O.encode_prequant(img, eY, eC, "f64")
    sym64 = O.quantise(pre64)
    tie = np.abs(pre64 * 255.0 - np.floor(pre64 * 255.0) - 0.5)
    lam = SYMBOL_MISMATCH_LIMIT * sym.size
    check_symbols(sym, sym64, tie, min_allow=int(lam + 4 * np.sqrt(lam) + 2))
    assert 0.02 < (sym64 > 0).mean() and np.abs(pre - pre64).max() < 5e-5
    d64 = O.decode_prequant(sym64, dY, dC, "f64")
    rec64 = np.round(d64 * 255.0).astype(np.uint8)
    tie_r = np.abs(d64 * 255.0 - np.floor(d64 * 255.0) - 0.5)
    lam = SYMBOL_MISMATCH_LIMIT * rec64.size
    check_symbols(dec(sym64), rec64, tie_r, min_allow=int(lam + 4 * np.sqrt(lam) + 2))
    assert enc.handle.saturated_activations() == 0 and dec.handle.saturated_activations() == 0
