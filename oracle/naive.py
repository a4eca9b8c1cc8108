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
