"""GPU tests of the forward extras of the reference's training step (run on a B200 with -m gpu):
Entropynet (tf2_0/src/training.py:25-42), the uniform-noise quantisation proxy (training.py:87-88) and tf.image.ssim
(training.py:108,113), each against the oracle, and the forward half of one training step composed from them.
Floating-point outputs: tolerances are stated at each comparison (the oracle's fp64 mode is the reference value)."""
import numpy as np
import pytest

from conftest import make_weights, synthetic_images
from oracle import nnic_oracle as O

pytestmark = pytest.mark.gpu


def _entropy_weights(nn, lh, lw, seed=21):
    w = nn.entropynet_glorot(lh, lw, seed, gain=1.0, bias_range=0.05)
    w["dense2/bias"] = np.array([3.0], np.float32)            # centre the output inside the clip range (0, 8)
    w["dense2/kernel"] = (w["dense2/kernel"] * 8).astype(np.float32)
    return w


def _encoded(nn, codec_factory, n, hh, ww, seed):
    """Real encoder outputs (float, before rounding) of n synthetic images, planes stacked on the batch axis."""
    enc, _ = codec_factory("spread", "tc_split")
    img = synthetic_images(n, hh, ww, seed=seed)
    planes = O.rgb_to_planes(img, "f32")
    return img, planes, np.concatenate(enc.run_model(planes), axis=0)


@pytest.mark.parametrize("shape", [(64, 128, 128), (1, 512, 768), (2, 120, 136), (3, 72, 40)])
def test_entropynet_against_oracle(nn, codec_factory, shape):
    """Config-3 patches (16 x 16 latent), one Kodak image (64 x 96 latent), an odd latent (15 x 17: FFMA convolutions) and a
    small ragged one; host and device buffers; clip at both ends."""
    import torch
    n, hh, ww = shape
    _img, _planes, x = _encoded(nn, codec_factory, n, hh, ww, seed=hh + ww)
    lh, lw = x.shape[1:3]
    w = _entropy_weights(nn, lh, lw)
    net = nn.Entropynet(0).set_weights(w)
    got = net(x)
    want = O.entropynet(x, w, "f64")
    assert got.shape == want.shape == (3 * n, 1) and got.dtype == np.float32
    assert want.min() > 0.01 and want.max() < 7.99, "test weights must keep the output inside the clip range"
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()
    assert np.abs(got - O.entropynet(x, w, "f32")).max() < 2e-4
    got_dev = net(torch.from_numpy(x).cuda())
    assert np.array_equal(got_dev.cpu().numpy(), got)
    # both clip ends (training.py:42)
    for bias, value in ((-50.0, 0.0), (50.0, 8.0)):
        w2 = dict(w); w2["dense2/bias"] = np.array([bias], np.float32)
        assert np.all(nn.Entropynet(0).set_weights(w2)(x[:2]) == value)
    with pytest.raises(nn.NnicError, match="features"):
        net(np.zeros((1, lh + 2, lw + 2, 32), np.float32))
    with pytest.raises(nn.NnicError, match="not set"):
        nn.Entropynet(0)(x[:1])


def test_noise_quantise(nn, codec_factory):
    """training.py:87-88.  With the uniform draw supplied the result is bit-identical to the fp32 restatement; the built-in
    generator is deterministic per seed, independent of host / device buffers, uniform in [-0.5, 0.5)/255 and clipped."""
    import torch
    enc, _ = codec_factory("spread", "tc_split")
    rng = np.random.default_rng(4)
    x = rng.random((3, 16, 24, 32)).astype(np.float32)
    x[0, 0, 0, :4] = [0.0, 1.0, 0.0005, 0.9995]
    u = (rng.random(x.shape) - 0.5).astype(np.float32)
    got = nn.noisy_quantise(enc.handle, x, noise=u)
    assert np.array_equal(got, O.noisy_quantise(x, u, "f32"))
    assert np.array_equal(nn.noisy_quantise(enc.handle, torch.from_numpy(x).cuda(), noise=torch.from_numpy(u).cuda()).cpu().numpy(), got)
    a = nn.noisy_quantise(enc.handle, x, seed=7)
    assert np.array_equal(a, nn.noisy_quantise(enc.handle, x, seed=7))
    assert np.array_equal(a, nn.noisy_quantise(enc.handle, torch.from_numpy(x).cuda(), seed=7).cpu().numpy())
    assert not np.array_equal(a, nn.noisy_quantise(enc.handle, x, seed=8))
    d = (a.astype(np.float64) - x) * 255.0
    inner = (x > 0.01) & (x < 0.99)
    assert a.min() >= 0.0 and a.max() <= 1.0 and np.abs(d[inner]).max() <= 0.5 + 1e-4
    assert abs(d[inner].mean()) < 0.01 and abs(d[inner].std() - np.sqrt(1 / 12)) < 0.01      # U(-0.5, 0.5)
    big = np.full((1 << 20) + 3, 0.5, np.float32)                                          # a count that is not a multiple of 4
    dd = (nn.noisy_quantise(enc.handle, big, seed=1).astype(np.float64) - 0.5) * 255.0
    hist, _ = np.histogram(dd, bins=16, range=(-0.5, 0.5))
    assert hist.min() > 0.9 * dd.size / 16 and hist.max() < 1.1 * dd.size / 16


@pytest.mark.parametrize("shape", [(2, 40, 56), (3, 128, 128), (1, 512, 768), (5, 11, 11), (2, 43, 75)])
def test_ssim_against_oracle(nn, codec_factory, shape):
    """tf.image.ssim(a, b, max_val=1.0): the separable fp32 evaluation on the GPU against the 121-tap fp64 oracle."""
    import torch
    enc, _ = codec_factory("spread", "tc_split")
    n, hh, ww = shape
    rng = np.random.default_rng(hh)
    a = (synthetic_images(n, hh, ww, seed=ww)[..., :1] / 255.0).astype(np.float32)
    b = np.clip(a + 0.08 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    got = nn.ssim(enc.handle, a, b)
    want = O.ssim(a, b, "f64")
    assert got.shape == (n,) and np.abs(got - want).max() < 5e-6, np.abs(got - want).max()
    assert np.abs(nn.ssim(enc.handle, a, a) - 1.0).max() < 1e-6
    assert np.array_equal(nn.ssim(enc.handle, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy(), got)
    assert np.array_equal(nn.ssim(enc.handle, a, b), got)                                   # deterministic reduction
    with pytest.raises(ValueError):
        nn.ssim(enc.handle, a[:, :10], b[:, :10])


def test_training_step_forward(nn, codec_factory):
    """The forward half of one training step of the reference (training.py:76-119, without the random flips and with the
    uniform draw fixed): planes -> BaseEncoder -> noise -> Entropynet / BaseDecoder -> SSIM -> the two losses, composed from
    the library's calls, against the same chain in the fp64 oracle."""
    eY, eC, dY, dC = make_weights("spread")
    enc, dec = codec_factory("spread", "tc_split")
    n, hh, ww = 4, 128, 128
    img = synthetic_images(n, hh, ww, seed=77)
    rng = np.random.default_rng(78)
    planes32 = O.rgb_to_planes(img, "f32")
    encoded = enc.run_model(planes32)
    batch_encoded = np.concatenate(encoded, axis=0)
    u = (rng.random(batch_encoded.shape) - 0.5).astype(np.float32)
    w = _entropy_weights(nn, hh // 8, ww // 8)
    net = nn.Entropynet(0).set_weights(w)
    noisy = nn.noisy_quantise(enc.handle, batch_encoded, noise=u)
    approx = net(batch_encoded)
    decoded = dec.run_model([noisy[0:n], noisy[n:2 * n], noisy[2 * n:]])
    ss = [nn.ssim(dec.handle, planes32[i], decoded[i]) for i in range(3)]
    coef = np.float32(0.01)
    loss_0 = (1 - ss[0].mean()) / 2 + coef * approx[:n]
    loss_1 = (1 - np.concatenate(ss[1:]).mean()) / 2 + np.float32(0.01) * approx[n:]
    # the same step in the fp64 oracle
    planes64 = O.rgb_to_planes(img, "f64")
    enc64 = [O.base_encoder(planes64[0], eY, "f64"), O.base_encoder(planes64[1], eC, "f64"), O.base_encoder(planes64[2], eC, "f64")]
    be64 = np.concatenate(enc64, axis=0)
    noisy64 = O.noisy_quantise(be64, u, "f64")
    approx64 = O.entropynet(be64, w, "f64")
    dec64 = [O.base_decoder(noisy64[0:n], dY, "f64"), O.base_decoder(noisy64[n:2 * n], dC, "f64"), O.base_decoder(noisy64[2 * n:], dC, "f64")]
    ss64 = [O.ssim(planes64[i], dec64[i], "f64") for i in range(3)]
    assert np.abs(batch_encoded - be64).max() < 2e-5
    assert np.abs(noisy - noisy64).max() < 2e-5
    assert np.abs(approx - approx64).max() < 5e-4
    for i in range(3):
        assert np.abs(decoded[i] - dec64[i]).max() < 5e-5
        assert np.abs(ss[i] - ss64[i]).max() < 5e-5
    want_0 = (1 - ss64[0].mean()) / 2 + 0.01 * approx64[:n]
    want_1 = (1 - np.concatenate(ss64[1:]).mean()) / 2 + 0.01 * approx64[n:]
    assert np.abs(loss_0 - want_0).max() < 1e-4 and np.abs(loss_1 - want_1).max() < 1e-4
