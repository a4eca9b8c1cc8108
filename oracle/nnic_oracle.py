"""CPU oracle for the tf2_0 codec hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under neural_network_image_compression_b200/
imports it, and the product path never falls back to it.

PARITY UNPINNED: the reference ships no golden vectors, no checkpoints and no value
assertions for this path, and its arithmetic lives in
TensorFlow (not installed here, un-vendored, version pinned only by the directory
names tf1_13/ and tf2_0/).  This file restates the published TensorFlow/Keras
semantics (SAME padding, Conv2D HWIO kernels, Conv2DTranspose HWOI kernels,
leaky_relu alpha=0.2, round-half-to-even) and anchors on the reference call sites
cited below.  It is pinned, as far as that is possible without TensorFlow, by two
implementations that share no code with it: the naive-loop NumPy convolutions of
oracle/naive.py and the plain-C restatement of the WHOLE path in oracle/nnic_oracle.c
(fp64; agrees to 5e-15 on even, odd and minimal sizes, tests/test_oracle.py).

Reference call sites restated here (paths relative to /root/reference):
  tf2_0/src/utils.py:7-9      colour constants (ycbcr_kernel, inverse via linalg.inv, offsets)
  tf2_0/src/utils.py:64-77    _project / convert_to_rgb / convert_to_colourspace
  tf2_0/src/utils.py:15-24    ProClass: models[0] for plane 0, models[1] for planes 1 and 2
  tf2_0/src/encoder.py:7-32   BaseEncoder (conv1, conv2, conv3, conv4, +res, conv8, clip)
  tf2_0/src/encoder.py:38-47  Encoder.__call__
  tf2_0/src/decoder.py:7-32   BaseDecoder (dconv1, dconv5, dconv6, +res, dconv7, dconv8, clip)
  tf2_0/src/decoder.py:39-48  Decoder.__call__
  tf1_13/src/training.py:62-71 discrete histogram + entropy

Two arithmetic modes:
  'f32'  mimics the reference dtype flow (every eager op rounded to fp32, conv
         accumulation order = whatever torch-CPU/oneDNN does, like TF's unspecified order)
  'f64'  the ideal result: same formulas, every op in fp64 (constants still take the
         values the fp32 flow would see only where the reference itself fixes them,
         i.e. nowhere: the colour matrices stay fp64 in this mode)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# --- tf2_0/src/utils.py:7-9 ---------------------------------------------------------
YCBCR_KERNEL = np.array([[0.299, 0.587, 0.114],
                         [-0.16874, -0.33126, 0.5],
                         [0.5, -0.41869, -0.08131]], dtype=np.float64)
YCBCR_INV_KERNEL = np.linalg.inv(YCBCR_KERNEL)
YCBCR_OFF = np.array([0.0, 0.5, 0.5], dtype=np.float64)

# (name, ksize, stride, cin, cout) -- tf2_0/src/encoder.py:10-17, decoder.py:10-17
ENCODER_LAYERS = (("conv1", 5, 2, 1, 32), ("conv2", 5, 2, 32, 64), ("conv3", 3, 1, 64, 64),
                  ("conv4", 3, 1, 64, 64), ("conv8", 5, 2, 64, 32))
DECODER_LAYERS = (("dconv1", 5, 2, 32, 64), ("dconv5", 3, 1, 64, 64), ("dconv6", 3, 1, 64, 64),
                  ("dconv7", 5, 2, 64, 64), ("dconv8", 5, 2, 64, 1))
LEAKY_ALPHA = 0.2  # tf.nn.leaky_relu default


def _np_dtype(mode):
    return {"f32": np.float32, "f64": np.float64}[mode]


def _t_dtype(mode):
    return {"f32": torch.float32, "f64": torch.float64}[mode]


def same_pad(in_size: int, k: int, s: int):
    """TensorFlow 'SAME': out=ceil(in/s); total=max((out-1)*s+k-in,0); before=total//2."""
    out = -(-in_size // s)
    tot = max((out - 1) * s + k - in_size, 0)
    return out, tot // 2, tot - tot // 2


def conv2d_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int, mode: str):
    """Keras Conv2D(padding='SAME'), NHWC input, kernel [kh,kw,Cin,Cout] + bias, no activation."""
    dt = _t_dtype(mode)
    kh, kw = kernel.shape[:2]
    _, pt, pb = same_pad(x.shape[1], kh, stride)
    _, pl, pr = same_pad(x.shape[2], kw, stride)
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dt).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(kernel)).to(dt).permute(3, 2, 0, 1).contiguous()
    bt = torch.from_numpy(np.ascontiguousarray(bias)).to(dt)
    y = F.conv2d(F.pad(xt, (pl, pr, pt, pb)), wt, None, stride=stride)
    y = y + bt.view(1, -1, 1, 1)  # BiasAdd is a separate rounded op in TF
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv2d_transpose_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int, mode: str):
    """Keras Conv2DTranspose(padding='SAME'), kernel [kh,kw,Cout,Cin]:
    full[s*i+a, s*j+b, co] += x[i,j,ci]*K[a,b,co,ci]; out = full[before:before+s*in]."""
    dt = _t_dtype(mode)
    kh, kw = kernel.shape[:2]
    H, W = x.shape[1], x.shape[2]
    _, pt, _ = same_pad(H * stride, kh, stride)
    _, pl, _ = same_pad(W * stride, kw, stride)
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dt).permute(0, 3, 1, 2)
    # torch conv_transpose2d weight layout is [Cin, Cout, kh, kw]
    wt = torch.from_numpy(np.ascontiguousarray(kernel)).to(dt).permute(3, 2, 0, 1).contiguous()
    bt = torch.from_numpy(np.ascontiguousarray(bias)).to(dt)
    full = F.conv_transpose2d(xt, wt, None, stride=stride)
    y = full[:, :, pt:pt + stride * H, pl:pl + stride * W]
    y = y + bt.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def leaky_relu(x: np.ndarray):
    a = x.dtype.type(LEAKY_ALPHA)
    return np.where(x > 0, x, x * a)


def base_encoder(x: np.ndarray, w: dict, mode: str):
    """tf2_0/src/encoder.py:19-32.  x: [P,H,W,1]; w: {'conv1/kernel':..., 'conv1/bias':...}."""
    def layer(name, k, s, v):
        return leaky_relu(conv2d_same(v, w[name + "/kernel"], w[name + "/bias"], s, mode))
    x = layer("conv1", 5, 2, x)
    x = layer("conv2", 5, 2, x)
    res = x
    x = layer("conv3", 3, 1, x)
    x = layer("conv4", 3, 1, x)
    x = x + res
    x = layer("conv8", 5, 2, x)
    return np.clip(x, 0, 1)


def base_decoder(x: np.ndarray, w: dict, mode: str):
    """tf2_0/src/decoder.py:19-32.  x: [P,h,w,32]."""
    def layer(name, k, s, v):
        return leaky_relu(conv2d_transpose_same(v, w[name + "/kernel"], w[name + "/bias"], s, mode))
    x = layer("dconv1", 5, 2, x)
    res = x
    x = layer("dconv5", 3, 1, x)
    x = layer("dconv6", 3, 1, x)
    x = x + res
    x = layer("dconv7", 5, 2, x)
    x = layer("dconv8", 5, 2, x)
    return np.clip(x, 0, 1)


def encoder_trace(x: np.ndarray, w: dict, mode: str):
    """Per-layer outputs of base_encoder: [conv1, conv2, conv3, conv4+res, clip(conv8)]."""
    def layer(name, s, v):
        return leaky_relu(conv2d_same(v, w[name + "/kernel"], w[name + "/bias"], s, mode))
    a1 = layer("conv1", 2, x)
    a2 = layer("conv2", 2, a1)
    a3 = layer("conv3", 1, a2)
    a4 = layer("conv4", 1, a3) + a2
    a5 = np.clip(layer("conv8", 2, a4), 0, 1)
    return [a1, a2, a3, a4, a5]


def decoder_trace(x: np.ndarray, w: dict, mode: str):
    """Per-layer outputs of base_decoder: [input, dconv1, dconv5, dconv6+res, dconv7, clip(dconv8)]."""
    def layer(name, s, v):
        return leaky_relu(conv2d_transpose_same(v, w[name + "/kernel"], w[name + "/bias"], s, mode))
    d1 = layer("dconv1", 2, x)
    d2 = layer("dconv5", 1, d1)
    d3 = layer("dconv6", 1, d2) + d1
    d4 = layer("dconv7", 2, d3)
    d5 = np.clip(layer("dconv8", 2, d4), 0, 1)
    return [x, d1, d2, d3, d4, d5]


def _project(kernel, t0, t1, t2):
    """tf2_0/src/utils.py:64-68: (t0*k0 + t1*k1) + t2*k2, each op rounded in the tensor dtype."""
    dt = t0.dtype.type
    outs = []
    for i in range(3):
        outs.append(t0 * dt(kernel[i, 0]) + t1 * dt(kernel[i, 1]) + t2 * dt(kernel[i, 2]))
    return outs


def rgb_to_planes(x_u8: np.ndarray, mode: str):
    """Encoder.__call__ lines 39-41: /255 then convert_to_colourspace.  Returns 3 x [N,H,W,1]."""
    dt = _np_dtype(mode)
    img = x_u8.astype(dt) / dt(255)
    t = [img[..., i:i + 1] for i in range(3)]
    o = _project(YCBCR_KERNEL, *t)
    return [o[i] + dt(YCBCR_OFF[i]) for i in range(3)]


def planes_to_rgb(planes, mode: str):
    """Decoder.__call__ lines 45-46: convert_to_rgb then clip(0,1).  planes: 3 x [N,H,W,1]."""
    dt = _np_dtype(mode)
    t = [np.asarray(planes[i], dtype=dt) - dt(YCBCR_OFF[i]) for i in range(3)]
    o = _project(YCBCR_INV_KERNEL, *t)
    return np.clip(np.concatenate(o, axis=3), 0, 1)


def encode_prequant(x_u8: np.ndarray, enc_y: dict, enc_cbcr: dict, mode: str = "f32"):
    """Encoder.__call__ up to (and including) the concat: float [N,h,w,96] in [0,1]."""
    planes = rgb_to_planes(x_u8, mode)
    outs = [base_encoder(planes[0], enc_y, mode), base_encoder(planes[1], enc_cbcr, mode),
            base_encoder(planes[2], enc_cbcr, mode)]
    return np.concatenate(outs, axis=3)


def quantise(prequant: np.ndarray):
    """encoder.py:47: np.round(encoded*255).astype(uint8) -- multiply in the tensor dtype, RN-even."""
    return np.round(prequant * prequant.dtype.type(255)).astype(np.uint8)


def encode(x_u8, enc_y, enc_cbcr, mode="f32"):
    return quantise(encode_prequant(x_u8, enc_y, enc_cbcr, mode))


def decode_planes(latent_u8: np.ndarray, dec_y: dict, dec_cbcr: dict, mode: str = "f32"):
    """decoder.py:40-43: /255, split in three 32-channel groups, three model calls."""
    dt = _np_dtype(mode)
    img = latent_u8.astype(dt) / dt(255)
    chans = [img[..., 32 * i:32 * (i + 1)] for i in range(3)]
    return [base_decoder(chans[0], dec_y, mode), base_decoder(chans[1], dec_cbcr, mode),
            base_decoder(chans[2], dec_cbcr, mode)]


def decode_prequant(latent_u8, dec_y, dec_cbcr, mode="f32"):
    """Float RGB in [0,1] before the final *255 / round."""
    return planes_to_rgb(decode_planes(latent_u8, dec_y, dec_cbcr, mode), mode)


def decode(latent_u8, dec_y, dec_cbcr, mode="f32"):
    d = decode_prequant(latent_u8, dec_y, dec_cbcr, mode)
    return np.round(d * d.dtype.type(255)).astype(np.uint8)


def histogram(latent_u8: np.ndarray):
    """tf1_13/src/training.py:62-68: per (image, plane) 256-bin counts.  -> int64 [N,3,256]."""
    n = latent_u8.shape[0]
    hist = np.zeros((n, 3, 256), dtype=np.int64)
    for i in range(n):
        for p in range(3):
            hist[i, p] = np.bincount(latent_u8[i, :, :, 32 * p:32 * (p + 1)].ravel(), minlength=256)
    return hist


def entropy_from_hist(hist: np.ndarray, mode: str = "f32"):
    """training.py:69-70: p=count/numel; H = sum p * (-log(clip(p,1e-5,1)) / log 2).  [..,256]->[..]."""
    dt = _np_dtype(mode)
    numel = hist.sum(axis=-1, keepdims=True).astype(dt)
    p = hist.astype(dt) / numel
    logof2 = np.log(dt(2))
    return (p * (-np.log(np.clip(p, dt(1e-5), dt(1.0))) / logof2)).sum(axis=-1, dtype=dt)


def rate(latent_u8: np.ndarray, H: int, W: int, mode: str = "f32"):
    """hist [N,3,256], entropy bits/symbol [N,3], bpp [N] (this build's definition, SURVEY 8a-a9:
    bpp_n = sum_p H_{n,p} * (h*w*32) / (H*W)), global hist [3,256]."""
    hist = histogram(latent_u8)
    ent = entropy_from_hist(hist, mode)
    h, w = latent_u8.shape[1:3]
    dt = _np_dtype(mode)
    bpp = (ent * dt(h * w * 32)).sum(axis=1, dtype=dt) / dt(H * W)
    return hist, ent, bpp, hist.sum(axis=0)


def pack_latent(latent_u8: np.ndarray):
    """utils.py:42-44 (`_feed_batch`, c == 96): picture[n, y, x, i] = byte number y*8w + x of the C-order stream of
    latent[n, :, :, 32i:32i+32].  Written as index arithmetic, independently of the product's reshape."""
    n, h, w, _ = latent_u8.shape
    flat = np.arange(4 * h * 8 * w).reshape(4 * h, 8 * w)
    pix, ch = flat // 32, flat % 32
    rows, cols = pix // w, pix % w
    out = np.empty((n, 4 * h, 8 * w, 3), np.uint8)
    for i in range(3):
        out[..., i] = latent_u8[:, rows, cols, 32 * i + ch]
    return out


def unpack_latent(picture_u8: np.ndarray):
    """utils.py:36-38 (`_feed_batch`, in_cshape == 96): the inverse mapping of pack_latent."""
    n, hh, ww, _ = picture_u8.shape
    h, w = hh // 4, ww // 8
    out = np.empty((n, h, w, 96), np.uint8)
    for i in range(3):
        stream = picture_u8[..., i].reshape(n, -1)
        for ch in range(32):
            out[..., 32 * i + ch] = stream[:, ch::32].reshape(n, h, w)
    return out


def psnr(a_u8: np.ndarray, b_u8: np.ndarray):
    mse = np.mean((a_u8.astype(np.float64) - b_u8.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


# ---- forward extras of the training step -----------------------------------------------------------
def entropynet(x: np.ndarray, w: dict, mode: str = "f32"):
    """tf2_0/src/training.py:25-42.  x [P,h,w,32] -> [P,1]: three SAME convolutions with leaky_relu, Flatten (NHWC order),
    Dense(512), Dense(1) (both linear), clip(0, 8)."""
    dt = _np_dtype(mode)
    v = np.asarray(x, dt)
    for name, s in (("conv1", 2), ("conv2", 1), ("conv3", 1)):
        v = leaky_relu(conv2d_same(v, w[name + "/kernel"], w[name + "/bias"], s, mode))
    v = v.reshape(v.shape[0], -1)
    v = v @ np.asarray(w["dense1/kernel"], dt) + np.asarray(w["dense1/bias"], dt)
    v = v @ np.asarray(w["dense2/kernel"], dt) + np.asarray(w["dense2/bias"], dt)
    return np.clip(v, 0, 8)


def noisy_quantise(encoded: np.ndarray, noise: np.ndarray, mode: str = "f32"):
    """training.py:87-88 with the uniform draw passed in: clip(encoded + noise / 255, 0, 1), noise in [-0.5, 0.5)."""
    dt = _np_dtype(mode)
    return np.clip(np.asarray(encoded, dt) + np.asarray(noise, dt) / dt(255), 0, 1)


def ssim(a: np.ndarray, b: np.ndarray, mode: str = "f32"):
    """tf.image.ssim(a, b, max_val=1.0) for [P,H,W,1] images (training.py:108,113), following TensorFlow's published
    definition (image_ops_impl._ssim_per_channel / _ssim_helper): an 11 x 11 Gaussian window (sigma 1.5, softmax-normalised),
    VALID depthwise filtering of x, y, x*y and x*x + y*y, c1 = (0.01)^2, c2 = (0.03)^2,
    luminance = (2 m0 m1 + c1) / (m0^2 + m1^2 + c1), cs = (2 E[xy] - 2 m0 m1 + c2) / (E[x^2 + y^2] - m0^2 - m1^2 + c2),
    ssim = mean over the map of luminance * cs.  The 2-D window is applied as written (121 taps), not separably."""
    dt = _t_dtype(mode)
    x = torch.from_numpy(np.ascontiguousarray(a)).to(dt).reshape(a.shape[0], 1, a.shape[1], a.shape[2])
    y = torch.from_numpy(np.ascontiguousarray(b)).to(dt).reshape(b.shape[0], 1, b.shape[1], b.shape[2])
    coords = torch.arange(11, dtype=dt) - 5.0
    g = -0.5 * coords ** 2 / (1.5 ** 2)
    k2d = torch.softmax((g.reshape(1, -1) + g.reshape(-1, 1)).reshape(-1), dim=0).reshape(1, 1, 11, 11)

    def red(t):
        return F.conv2d(t, k2d)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m0, m1 = red(x), red(y)
    num0 = m0 * m1 * 2.0
    den0 = m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    num1 = red(x * y) * 2.0
    den1 = red(x * x + y * y)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(dim=(1, 2, 3)).numpy()
