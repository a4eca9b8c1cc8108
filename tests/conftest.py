import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SETS = {"default": (1.0, 0.0), "spread": (1.6, 0.05)}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def make_weights(name):
    """(encY, encCbCr, decY, decCbCr) for a named weight set -- same recipe as tools/make_golden.py."""
    from neural_network_image_compression_b200 import weights as Wt
    gain, br = WEIGHT_SETS[name]
    return (Wt.glorot_uniform("encoder", 11, gain, br), Wt.glorot_uniform("encoder", 12, gain, br),
            Wt.glorot_uniform("decoder", 13, gain, br), Wt.glorot_uniform("decoder", 14, gain, br))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def synthetic_images(n, h, w, seed=0):
    """Blocky, noisy RGB images (natural-ish statistics, all 256 levels present)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(n, -(-h // 4), -(-w // 4), 3)).astype(np.float32)
    up = np.kron(base, np.ones((1, 4, 4, 1), np.float32))[:, :h, :w]
    return np.clip(up + rng.normal(0, 12, size=(n, h, w, 3)), 0, 255).astype(np.uint8)


@pytest.fixture(scope="session")
def nn():
    import neural_network_image_compression_b200 as nn_
    return nn_


@pytest.fixture(scope="session")
def codec_factory(nn):
    """codec_factory(weight_set, arith) -> (Encoder, Decoder), cached for the session."""
    cache = {}

    def get(wname="default", arith="tc_split"):
        key = (wname, arith)
        if key not in cache:
            eY, eC, dY, dC = make_weights(wname)
            enc, dec = nn.Encoder(0, arith), nn.Decoder(0, arith)
            enc.set_weights(0, eY); enc.set_weights(1, eC)
            dec.set_weights(0, dY); dec.set_weights(1, dC)
            cache[key] = (enc, dec)
        return cache[key]
    return get
