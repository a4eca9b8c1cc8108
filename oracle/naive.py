"""Naive-loop restatement of TensorFlow's SAME Conv2D / Conv2DTranspose.  TEST INFRASTRUCTURE ONLY.

Written straight from the TensorFlow definitions, with no torch, so that oracle/nnic_oracle.py's
torch mapping (pad + conv2d, conv_transpose2d + crop) is checked by an independent formulation.
Pure Python/NumPy loops: use on small shapes only.

  Conv2D (tf2_0/src/encoder.py:10-17):        out[n,y,x,co] = b[co] + sum_{a,b,ci} xpad[n, y*s+a, x*s+b, ci] * K[a,b,ci,co]
  Conv2DTranspose (tf2_0/src/decoder.py:10-17): full[n, i*s+a, j*s+b, co] += x[n,i,j,ci] * K[a,b,co,ci]
                                               out = full[:, pt:pt+s*H, pl:pl+s*W] + b
"""
import numpy as np


def same_pad(in_size, k, s):
    out = -(-in_size // s)
    tot = max((out - 1) * s + k - in_size, 0)
    return out, tot // 2, tot - tot // 2


def conv2d_same_naive(x, kernel, bias, stride):
    x = np.asarray(x, np.float64)
    kernel = np.asarray(kernel, np.float64)
    n, H, W, cin = x.shape
    kh, kw, _, cout = kernel.shape
    oh, pt, pb = same_pad(H, kh, stride)
    ow, pl, pr = same_pad(W, kw, stride)
    xp = np.zeros((n, H + pt + pb, W + pl + pr, cin))
    xp[:, pt:pt + H, pl:pl + W] = x
    out = np.zeros((n, oh, ow, cout))
    for y in range(oh):
        for xx in range(ow):
            patch = xp[:, y * stride:y * stride + kh, xx * stride:xx * stride + kw, :]
            out[:, y, xx, :] = np.tensordot(patch, kernel, axes=([1, 2, 3], [0, 1, 2]))
    return out + np.asarray(bias, np.float64)


def conv2d_transpose_same_naive(x, kernel, bias, stride):
    x = np.asarray(x, np.float64)
    kernel = np.asarray(kernel, np.float64)
    n, H, W, cin = x.shape
    kh, kw, cout, _ = kernel.shape
    full = np.zeros((n, (H - 1) * stride + kh, (W - 1) * stride + kw, cout))
    for i in range(H):
        for j in range(W):
            # contribution[n,a,b,co] = sum_ci x[n,i,j,ci] * K[a,b,co,ci]
            full[:, i * stride:i * stride + kh, j * stride:j * stride + kw, :] += np.einsum(
                "nc,aboc->nabo", x[:, i, j, :], kernel)
    _, pt, _ = same_pad(H * stride, kh, stride)
    _, pl, _ = same_pad(W * stride, kw, stride)
    out = full[:, pt:pt + stride * H, pl:pl + stride * W, :]
    return out + np.asarray(bias, np.float64)


def dconv8_by_tap_responses(x, kernel, bias, tile_rows=16, tile_cols=8):
    """Conv2DTranspose(1, 5, 2, 'SAME') (reference decoder.py:17) restated the way the CUDA decoder tail computes it
    (csrc/tc_conv_patch.cu FUSE8 + csrc/tc_dconv8.cu k_dconv8_gather): x [n, 2Hp, 2Wp, Cin] is dconv7's output, addressed
    as (y, x) = (2 iy + py, 2 ix + px) over dconv7's Hp x Wp input grid;
      R[n][tile][t = 5a + b][phase = 2 py + px][m] = sum_ci x[n, y, x, ci] * K[a, b, 0, ci]
    with tiles of tile_rows x tile_cols input-grid pixels in row-major order and m = tile_cols * (iy % tile_rows) + ix % tile_cols,
    and every response feeds exactly one output pixel: out[2y + a - 1, 2x + b - 1] += R[y, x, 5a + b].
    Returns (out [n, 4Hp, 4Wp, 1] before the activation, R).  Test infrastructure only."""
    x = np.asarray(x, np.float64)
    kernel = np.asarray(kernel, np.float64)
    n, H2, W2, cin = x.shape
    assert H2 % 2 == 0 and W2 % 2 == 0 and kernel.shape == (5, 5, 1, cin)
    Hp, Wp = H2 // 2, W2 // 2
    ty, tx = -(-Hp // tile_rows), -(-Wp // tile_cols)
    R = np.zeros((n, ty * tx, 25, 4, tile_rows * tile_cols))
    for iy in range(Hp):
        for ix in range(Wp):
            tile = (iy // tile_rows) * tx + ix // tile_cols
            m = tile_cols * (iy % tile_rows) + ix % tile_cols
            for py in range(2):
                for px in range(2):
                    v = x[:, 2 * iy + py, 2 * ix + px, :]                       # [n, cin]
                    R[:, tile, :, 2 * py + px, m] = np.einsum("nc,abc->nab", v, kernel[:, :, 0, :]).reshape(n, 25)
    out = np.zeros((n, 2 * H2, 2 * W2, 1))
    for oy in range(2 * H2):
        for ox in range(2 * W2):
            acc = np.zeros(n)
            for a in range((oy + 1) & 1, 5, 2):                                 # 2y + a - 1 = oy
                y = (oy + 1 - a) // 2
                if not 0 <= y < H2:
                    continue
                for b in range((ox + 1) & 1, 5, 2):
                    xx = (ox + 1 - b) // 2
                    if not 0 <= xx < W2:
                        continue
                    iy, py, ix, px = y // 2, y & 1, xx // 2, xx & 1
                    tile = (iy // tile_rows) * tx + ix // tile_cols
                    m = tile_cols * (iy % tile_rows) + ix % tile_cols
                    acc += R[:, tile, 5 * a + b, 2 * py + px, m]
            out[:, oy, ox, 0] = acc
    return out + np.asarray(bias, np.float64), R
