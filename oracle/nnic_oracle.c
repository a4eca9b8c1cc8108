/* Plain-C restatement of the tf2_0 codec hot path.  TEST INFRASTRUCTURE ONLY (built by oracle/Makefile into
 * oracle/_build/libnnic_oracle_c.so, loaded by tests/test_oracle.py); nothing under neural_network_image_compression_b200/
 * uses it.
 *
 * It exists to pin oracle/nnic_oracle.py (torch-CPU convolutions + NumPy) against a second implementation that shares no
 * code with it: every layer below is the textbook definition with explicit loops, written from the reference call sites
 *   tf2_0/src/encoder.py:7-32,38-47   BaseEncoder / Encoder.__call__
 *   tf2_0/src/decoder.py:7-32,39-48   BaseDecoder / Decoder.__call__
 *   tf2_0/src/utils.py:7-9,64-77      colour constants and projections
 * and the published TensorFlow semantics (SAME padding, Conv2D HWIO kernels, Conv2DTranspose HWOI kernels, leaky_relu
 * alpha 0.2, round-half-to-even).  PARITY UNPINNED against TensorFlow itself, like the Python oracle.
 *
 * All arithmetic is in `real` (double by default: the "ideal" result the tie band of the parity tests is defined on).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

static const double K_YCBCR[3][3] = {{0.299, 0.587, 0.114}, {-0.16874, -0.33126, 0.5}, {0.5, -0.41869, -0.08131}};
static const double OFF_YCBCR[3] = {0.0, 0.5, 0.5};

static real leaky(real v) { return v > 0 ? v : (real)0.2 * v; }
static real clip01(real v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }

/* TF SAME: out = ceil(in / s), total padding = max((out-1)*s + k - in, 0), before = total / 2 */
static void same_pad(int in, int k, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int tot = (*out - 1) * s + k - in;
  if (tot < 0) tot = 0;
  *before = tot / 2;
}

/* Conv2D(cout, k, s, 'SAME') + bias + leaky_relu.  x [H,W,Cin], w [k,k,Cin,Cout] (HWIO), y [Ho,Wo,Cout] */
static real* conv2d(const real* x, int H, int W, int cin, const float* w, const float* b, int k, int s, int cout,
                    int* Ho, int* Wo) {
  int pt, pl;
  same_pad(H, k, s, Ho, &pt);
  same_pad(W, k, s, Wo, &pl);
  real* y = (real*)malloc(sizeof(real) * (size_t)(*Ho) * (*Wo) * cout);
  for (int oy = 0; oy < *Ho; ++oy)
    for (int ox = 0; ox < *Wo; ++ox)
      for (int co = 0; co < cout; ++co) {
        real acc = 0;
        for (int a = 0; a < k; ++a) {
          const int iy = oy * s - pt + a;
          if (iy < 0 || iy >= H) continue;
          for (int c = 0; c < k; ++c) {
            const int ix = ox * s - pl + c;
            if (ix < 0 || ix >= W) continue;
            const real* px = x + ((size_t)iy * W + ix) * cin;
            const float* wk = w + ((size_t)(a * k + c) * cin) * cout + co;
            for (int ci = 0; ci < cin; ++ci) acc += px[ci] * (real)wk[(size_t)ci * cout];
          }
        }
        y[((size_t)oy * (*Wo) + ox) * cout + co] = leaky(acc + (real)b[co]);
      }
  return y;
}

/* Conv2DTranspose(cout, k, s, 'SAME') + bias + leaky_relu.  w [k,k,Cout,Cin] (HWOI).  Scatter form:
 * full[i*s + a, j*s + c, co] += x[i,j,ci] * w[a,c,co,ci]; the output is full[before : before + H*s] per axis. */
static real* conv2d_transpose(const real* x, int H, int W, int cin, const float* w, const float* b, int k, int s, int cout,
                              int* Ho, int* Wo) {
  *Ho = H * s; *Wo = W * s;
  int o_, pt, pl;
  same_pad(*Ho, k, s, &o_, &pt);
  same_pad(*Wo, k, s, &o_, &pl);
  real* y = (real*)calloc((size_t)(*Ho) * (*Wo) * cout, sizeof(real));
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j)
      for (int a = 0; a < k; ++a) {
        const int oy = i * s + a - pt;
        if (oy < 0 || oy >= *Ho) continue;
        for (int c = 0; c < k; ++c) {
          const int ox = j * s + c - pl;
          if (ox < 0 || ox >= *Wo) continue;
          const real* px = x + ((size_t)i * W + j) * cin;
          for (int co = 0; co < cout; ++co) {
            const float* wk = w + ((size_t)(a * k + c) * cout + co) * cin;
            real acc = 0;
            for (int ci = 0; ci < cin; ++ci) acc += px[ci] * (real)wk[ci];
            y[((size_t)oy * (*Wo) + ox) * cout + co] += acc;
          }
        }
      }
  for (size_t i = 0; i < (size_t)(*Ho) * (*Wo); ++i)
    for (int co = 0; co < cout; ++co) y[i * cout + co] = leaky(y[i * cout + co] + (real)b[co]);
  return y;
}

/* weights of one network: 5 (kernel, bias) pairs in call order */
typedef struct { const float* k[5]; const float* b[5]; } net_t;

/* BaseEncoder.call (encoder.py:19-32): one plane [H,W,1] -> [h,w,32] clipped */
static real* base_encoder(const real* plane, int H, int W, const net_t* n, int* h, int* w) {
  int H1, W1, H2, W2, t0, t1;
  real* a1 = conv2d(plane, H, W, 1, n->k[0], n->b[0], 5, 2, 32, &H1, &W1);
  real* a2 = conv2d(a1, H1, W1, 32, n->k[1], n->b[1], 5, 2, 64, &H2, &W2);
  real* a3 = conv2d(a2, H2, W2, 64, n->k[2], n->b[2], 3, 1, 64, &t0, &t1);
  real* a4 = conv2d(a3, H2, W2, 64, n->k[3], n->b[3], 3, 1, 64, &t0, &t1);
  for (size_t i = 0; i < (size_t)H2 * W2 * 64; ++i) a4[i] += a2[i];
  real* a5 = conv2d(a4, H2, W2, 64, n->k[4], n->b[4], 5, 2, 32, h, w);
  for (size_t i = 0; i < (size_t)(*h) * (*w) * 32; ++i) a5[i] = clip01(a5[i]);
  free(a1); free(a2); free(a3); free(a4);
  return a5;
}

/* BaseDecoder.call (decoder.py:19-32): [h,w,32] -> [8h,8w,1] clipped */
static real* base_decoder(const real* x, int h, int w, const net_t* n) {
  int H1, W1, H2, W2, H3, W3, t0, t1;
  real* d1 = conv2d_transpose(x, h, w, 32, n->k[0], n->b[0], 5, 2, 64, &H1, &W1);
  real* d2 = conv2d_transpose(d1, H1, W1, 64, n->k[1], n->b[1], 3, 1, 64, &t0, &t1);
  real* d3 = conv2d_transpose(d2, H1, W1, 64, n->k[2], n->b[2], 3, 1, 64, &t0, &t1);
  for (size_t i = 0; i < (size_t)H1 * W1 * 64; ++i) d3[i] += d1[i];
  real* d4 = conv2d_transpose(d3, H1, W1, 64, n->k[3], n->b[3], 5, 2, 64, &H2, &W2);
  real* d5 = conv2d_transpose(d4, H2, W2, 64, n->k[4], n->b[4], 5, 2, 1, &H3, &W3);
  for (size_t i = 0; i < (size_t)H3 * W3; ++i) d5[i] = clip01(d5[i]);
  free(d1); free(d2); free(d3); free(d4);
  return d5;
}

static void fill_net(net_t* n, const float* const* ptrs) {
  for (int l = 0; l < 5; ++l) { n->k[l] = ptrs[2 * l]; n->b[l] = ptrs[2 * l + 1]; }
}

/* Encoder.__call__ up to the concat (encoder.py:39-45): rgb u8 [H,W,3] -> prequant double [h,w,96].
 * w_y / w_cbcr: 10 pointers each (kernel, bias per layer).  Returns 0, or -1 on a size mismatch. */
int oracle_c_encode_prequant(const uint8_t* rgb, int H, int W, const float* const* w_y, const float* const* w_cbcr,
                             double* out, int h_expect, int w_expect) {
  net_t ny, nc;
  fill_net(&ny, w_y); fill_net(&nc, w_cbcr);
  real* plane = (real*)malloc(sizeof(real) * (size_t)H * W);
  for (int p = 0; p < 3; ++p) {
    for (size_t i = 0; i < (size_t)H * W; ++i) {
      const real r = (real)rgb[3 * i] / 255, g = (real)rgb[3 * i + 1] / 255, b = (real)rgb[3 * i + 2] / 255;
      plane[i] = (r * (real)K_YCBCR[p][0] + g * (real)K_YCBCR[p][1]) + b * (real)K_YCBCR[p][2] + (real)OFF_YCBCR[p];
    }
    int h, w;
    real* e = base_encoder(plane, H, W, p == 0 ? &ny : &nc, &h, &w);
    if (h != h_expect || w != w_expect) { free(e); free(plane); return -1; }
    for (size_t i = 0; i < (size_t)h * w; ++i)
      for (int c = 0; c < 32; ++c) out[i * 96 + 32 * p + c] = (double)e[i * 32 + c];
    free(e);
  }
  free(plane);
  return 0;
}

/* 3x3 inverse by cofactors (np.linalg.inv of ycbcr_kernel, utils.py:8) */
static void inv3(const double m[3][3], double o[3][3]) {
  const double d = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                   m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  o[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / d; o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / d;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / d; o[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / d;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / d; o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / d;
  o[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / d; o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / d;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / d;
}

/* Decoder.__call__ up to the clip (decoder.py:40-46): latent u8 [h,w,96] -> rgb in [0,1], double [8h,8w,3] */
int oracle_c_decode_prequant(const uint8_t* latent, int h, int w, const float* const* w_y, const float* const* w_cbcr,
                             double* out) {
  net_t ny, nc;
  fill_net(&ny, w_y); fill_net(&nc, w_cbcr);
  const size_t npx = (size_t)64 * h * w;
  real* planes[3];
  real* x = (real*)malloc(sizeof(real) * (size_t)h * w * 32);
  for (int p = 0; p < 3; ++p) {
    for (size_t i = 0; i < (size_t)h * w; ++i)
      for (int c = 0; c < 32; ++c) x[i * 32 + c] = (real)latent[i * 96 + 32 * p + c] / 255;
    planes[p] = base_decoder(x, h, w, p == 0 ? &ny : &nc);
  }
  free(x);
  double kinv[3][3];
  inv3(K_YCBCR, kinv);
  for (size_t i = 0; i < npx; ++i) {
    const real t0 = planes[0][i] - (real)OFF_YCBCR[0], t1 = planes[1][i] - (real)OFF_YCBCR[1], t2 = planes[2][i] - (real)OFF_YCBCR[2];
    for (int k = 0; k < 3; ++k)
      out[3 * i + k] = (double)clip01((t0 * (real)kinv[k][0] + t1 * (real)kinv[k][1]) + t2 * (real)kinv[k][2]);
  }
  for (int p = 0; p < 3; ++p) free(planes[p]);
  return 0;
}

/* np.round(v * 255).astype(uint8) with round-half-to-even (encoder.py:47, decoder.py:48) */
void oracle_c_quantise(const double* v, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)nearbyint(v[i] * 255.0);
}
