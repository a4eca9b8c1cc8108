"""CPU tests of the oracle: TensorFlow semantics pinned by an independent naive restatement, scalar
semantics, and the committed golden vectors."""
import numpy as np
import pytest

from conftest import load_golden, make_weights
from oracle import naive
from oracle import nnic_oracle as O


@pytest.mark.parametrize("hw", [(8, 12), (7, 9), (5, 5), (1, 3)])
@pytest.mark.parametrize("ks", [(5, 2), (3, 1)])
def test_conv_matches_naive(hw, ks):
    rng = np.random.default_rng(1)
    k, s = ks
    x = rng.standard_normal((2, hw[0], hw[1], 3))
    K = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    got = O.conv2d_same(x, K, b, s, "f64")
    ref = naive.conv2d_same_naive(x, K, b, s)
    assert got.shape == ref.shape == (2, -(-hw[0] // s), -(-hw[1] // s), 4)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("hw", [(8, 12), (7, 9), (1, 1)])
@pytest.mark.parametrize("ks", [(5, 2), (3, 1)])
def test_conv_transpose_matches_naive(hw, ks):
    rng = np.random.default_rng(2)
    k, s = ks
    x = rng.standard_normal((2, hw[0], hw[1], 3))
    K = rng.standard_normal((k, k, 4, 3))     # [kh,kw,Cout,Cin]
    b = rng.standard_normal(4)
    got = O.conv2d_transpose_same(x, K, b, s, "f64")
    ref = naive.conv2d_transpose_same_naive(x, K, b, s)
    assert got.shape == ref.shape == (2, hw[0] * s, hw[1] * s, 4)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("hw", [(2, 2), (6, 10), (34, 18)])
def test_dconv8_as_tap_responses_plus_gather(hw):
    """The decoder tail of the CUDA path (dconv7's epilogue computes one response per (dconv7 output pixel, dconv8 tap), stored
    tile-blocked; a gather kernel adds the responses that reach each output pixel) is the reference's
    Conv2DTranspose(1, 5, 2, 'SAME') (decoder.py:17): checked here on the CPU against the two restatements of that layer,
    including a grid that spans several 16 x 8 tiles with ragged edges."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, hw[0], hw[1], 5))
    K = rng.standard_normal((5, 5, 1, 5))     # [kh,kw,Cout,Cin]
    b = rng.standard_normal(1)
    got, R = naive.dconv8_by_tap_responses(x, K, b)
    assert got.shape == (2, 2 * hw[0], 2 * hw[1], 1)
    assert np.abs(got - naive.conv2d_transpose_same_naive(x, K, b, 2)).max() < 1e-12
    assert np.abs(got - O.conv2d_transpose_same(x, K, b, 2, "f64")).max() < 1e-12
    # every response is used exactly once, except those whose output pixel lies outside the image (SAME cropping)
    tiles = -(-(hw[0] // 2) // 16) * -(-(hw[1] // 2) // 8)
    assert R.shape == (2, tiles, 25, 4, 128)


def test_same_padding_rule():
    # k5 s2: even sizes pad (1,2), odd sizes (2,2); k3 s1: (1,1)
    assert O.same_pad(8, 5, 2) == (4, 1, 2)
    assert O.same_pad(7, 5, 2) == (4, 2, 2)
    assert O.same_pad(8, 3, 1) == (8, 1, 1)
    assert O.same_pad(1, 5, 2) == (1, 2, 2)


def test_scalar_semantics():
    # true division by 255 differs from multiplying by the fp32 reciprocal for many byte values
    v = np.arange(256, dtype=np.float32)
    assert int(np.sum(v / np.float32(255) != v * np.float32(1 / 255))) > 100
    # round half to even
    assert O.quantise(np.array([0.5 / 255, 1.5 / 255, 2.5 / 255], np.float64)).tolist() == [0, 2, 2]
    # leaky relu: alpha 0.2 on the negative side only
    assert np.allclose(O.leaky_relu(np.array([-1.0, 2.0], np.float32)), [-0.2, 2.0])


def test_colour_roundtrip_and_constants():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(1, 6, 5, 3), dtype=np.uint8)
    planes = O.rgb_to_planes(img, "f64")
    back = O.planes_to_rgb(planes, "f64")
    assert np.abs(back * 255 - img).max() < 1e-9
    # fp32 flow keeps fp32 everywhere
    assert all(p.dtype == np.float32 for p in O.rgb_to_planes(img, "f32"))
    assert O.planes_to_rgb(O.rgb_to_planes(img, "f32"), "f32").dtype == np.float32
    # Y row sums to 1, chroma rows to 0
    assert abs(O.YCBCR_KERNEL[0].sum() - 1) < 1e-12 and abs(O.YCBCR_KERNEL[1].sum()) < 1e-12


def test_histogram_entropy_definition():
    rng = np.random.default_rng(4)
    lat = rng.integers(0, 7, size=(2, 3, 4, 96), dtype=np.uint8)
    lat[0, :, :, :32] = 5                      # a constant plane: one bin, entropy 0
    hist, ent, bpp, hg = O.rate(lat, 24, 32, "f64")
    assert hist.shape == (2, 3, 256) and hist.sum() == lat.size
    assert hist[0, 0, 5] == 3 * 4 * 32 and ent[0, 0] == 0
    p = hist[1, 2] / hist[1, 2].sum()
    direct = -(p[p > 0] * np.log2(np.maximum(p[p > 0], 1e-5))).sum()
    assert abs(ent[1, 2] - direct) < 1e-12
    assert np.allclose(bpp, ent.sum(axis=1) * (3 * 4 * 32) / (24 * 32))
    assert np.array_equal(hg, hist.sum(axis=0))
    # the clip at 1e-5: a bin with probability below 1e-5 contributes p*log2(1e5), not p*log2(1/p)
    h = np.zeros((1, 256), np.int64); h[0, 0] = 10 ** 6; h[0, 1] = 1
    e = O.entropy_from_hist(h, "f64")[0]
    p1 = 1 / (10 ** 6 + 1)
    assert abs(e - (p1 * np.log2(1e5) - (1 - p1) * np.log2(1 - p1))) < 1e-12


def test_shapes_ragged_and_empty_batch():
    eY, eC, dY, dC = make_weights("default")
    img = np.zeros((1, 20, 13, 3), np.uint8)           # not a multiple of 8: ceil at every stride-2 stage
    sym = O.encode(img, eY, eC, "f32")
    assert sym.shape == (1, 3, 2, 96) and sym.dtype == np.uint8
    rec = O.decode(sym, dY, dC, "f32")
    assert rec.shape == (1, 24, 16, 3)                 # 8*h, not cropped (decoder.py:39-48)


@pytest.mark.parametrize("name", ["kodim21_crop", "imagenet_patches"])
@pytest.mark.parametrize("wname", ["default", "spread"])
def test_golden_vectors(name, wname):
    g = load_golden(name)
    eY, eC, dY, dC = make_weights(wname)
    img = g["input"]
    # fp64 mode reproduces the stored symbols exactly (a flip would need a value within 1e-13 of a tie)
    sym64 = O.encode(img, eY, eC, "f64")
    assert np.array_equal(sym64, g[f"{wname}_sym64"])
    # fp32 mode: summation order may differ between CPUs -> only rounding ties may flip
    sym32 = O.encode(img, eY, eC, "f32")
    diff = sym32 != g[f"{wname}_sym64"]
    assert diff.mean() <= 1e-4
    assert np.abs(sym32.astype(int) - g[f"{wname}_sym64"].astype(int)).max() <= 1
    assert np.all(g[f"{wname}_tie_dist"][diff] < 2e-3)
    rec64 = O.decode(g[f"{wname}_sym64"], dY, dC, "f64")
    assert np.array_equal(rec64, g[f"{wname}_rec64"])
    hist, ent, bpp, _ = O.rate(g[f"{wname}_sym64"], img.shape[1], img.shape[2], "f32")
    assert np.array_equal(hist, g[f"{wname}_hist"].astype(np.int64))
    assert np.allclose(bpp, g[f"{wname}_bpp"], rtol=1e-5)


# ---- second, independent restatement in plain C (oracle/nnic_oracle.c) ---------------------------------------------
def test_c_restatement_agrees_with_the_python_oracle():
    """oracle/nnic_oracle.c (explicit loops, written from the reference call sites, shares no code with the torch-based
    oracle) must give the same fp64 pre-quantisation values to rounding noise, for even, odd and minimal sizes, both
    weight sets, encoder and decoder; hence identical symbols / bytes except exactly at ties."""
    from conftest import make_weights, synthetic_images, load_golden
    from oracle import c_oracle as CO
    for wname in ("default", "spread"):
        eY, eC, dY, dC = make_weights(wname)
        cases = [synthetic_images(2, 40, 56, seed=5), synthetic_images(1, 37, 29, seed=6), synthetic_images(1, 8, 8, seed=7),
                 load_golden("kodim21_crop")["input"][:, :64, :96]]
        for img in cases:
            a = CO.encode_prequant(img, eY, eC)
            b = O.encode_prequant(img, eY, eC, "f64")
            assert a.shape == b.shape and np.abs(a - b).max() < 1e-12
            sym = O.quantise(b)
            tie = np.abs(b * 255.0 - np.floor(b * 255.0) - 0.5)
            diff = CO.quantise(a) != sym
            assert np.all(tie[diff] < 1e-9)
            assert np.array_equal(CO.quantise(b), sym)                   # round-half-to-even in C == np.round
            ra = CO.decode_prequant(sym, dY, dC)
            rb = O.decode_prequant(sym, dY, dC, "f64")
            assert ra.shape == rb.shape and np.abs(ra - rb).max() < 1e-12


# ---- the full-size parity records (tests/parity_fixture.py) describe what the oracle computes today ----------------
@pytest.mark.parametrize("config,wname,indices", [("c2", "spread", (0, 17)), ("c2", "default", (5,)),
                                                  ("c3", "spread", (0, 1, 2047, 4095)), ("c3", "default", (7, 3000)),
                                                  ("c5", "spread", (0, 2047)), ("c5", "default", (1024,))])
def test_parity_records_against_live_oracle(config, wname, indices):
    """Sampled images of every full-size record are recomputed with the live oracle (fp64 and fp32): same digest of the
    bytes outside the tie band, same tie positions and tie bytes.  (Config 4's 4K images take minutes per image in fp64;
    its record is produced by the same code path, tools/make_parity_fixtures.py::make.)"""
    import parity_fixture as PF
    rec = PF.load_record(f"parity_{config}_{wname}")
    eY, eC, dY, dC = make_weights(wname)
    if config == "c2":
        images = PF.c2_images()
    elif config == "c3":
        images = PF.c3_patches()
    else:
        images = np.concatenate([PF.c5_patches(i, 1).numpy() for i in indices])
    assert tuple(rec["shape"][1:]) == images.shape[1:]
    img = images[list(indices)] if config != "c5" else images
    pre64 = O.encode_prequant(img, eY, eC, "f64")
    PF.verify_images_with_live_oracle(rec, "lat", indices, pre64 * 255.0, O.encode(img, eY, eC, "f32"))
    if config == "c2":
        sym64 = O.quantise(pre64)
        d64 = O.decode_prequant(sym64, dY, dC, "f64")
        PF.verify_images_with_live_oracle(rec, "rec", indices, d64 * 255.0, O.decode(sym64, dY, dC, "f32"))


def test_parity_record_compare_detects_a_wrong_byte():
    """compare() must reject one flipped byte outside the tie band and a tie byte that is neither floor nor floor + 1."""
    import parity_fixture as PF
    rec = PF.load_record("parity_c2_spread")
    eY, eC, _dY, _dC = make_weights("spread")
    img = PF.c2_images()[:1]
    sym = O.encode(img, eY, eC, "f64")
    st = PF.compare(sym, rec, "lat", first_image=0)
    assert st["mismatch_vs_f64"] == 0 and st["mismatch_outside_tie_band"] == 0
    ties = set(PF.tie_positions(rec, "lat")[:100000].tolist())
    k = next(i for i in range(sym.size) if i not in ties)
    bad = sym.copy(); bad.reshape(-1)[k] ^= 1
    with pytest.raises(AssertionError, match="outside the rounding-tie band"):
        PF.compare(bad, rec, "lat", first_image=0)
    t = int(PF.tie_positions(rec, "lat")[0])
    bad = sym.copy(); bad.reshape(-1)[t] = (int(rec["lat_tie_floor"][0]) + 2) & 0xff
    with pytest.raises(AssertionError, match="neither floor nor floor"):
        PF.compare(bad, rec, "lat", first_image=0)
    # a legal tie flip is counted, not rejected
    ok = sym.copy(); f = int(rec["lat_tie_floor"][0]); ok.reshape(-1)[t] = f + 1 if int(sym.reshape(-1)[t]) == f else f
    assert PF.compare(ok, rec, "lat", first_image=0)["mismatch_vs_f64"] == 1
