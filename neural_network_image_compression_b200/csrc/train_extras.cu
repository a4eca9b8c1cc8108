// Forward-only kernels of the reference's training step that lie beside the codec path:
//   * the dense head of Entropynet (tf2_0/src/training.py:31-32,39-42: Flatten, Dense(512), Dense(1), clip(0, 8)); its three
//     convolutions run on the tensor-core convolution kernel (same layer types as conv2 / conv3);
//   * the uniform-noise quantisation proxy (training.py:87-88): clip(encoded + U(-0.5, 0.5) / 255, 0, 1);
//   * tf.image.ssim(a, b, max_val=1.0) of single-channel images (training.py:108,113): 11 x 11 Gaussian window, sigma 1.5,
//     k1 = 0.01, k2 = 0.03, VALID windows, mean over the (H-10) x (W-10) map.
// All fp32, HBM-bound except the dense layer (an fp32 FFMA GEMM: 2 * F * 512 flops per plane, F = 64 * h/2 * w/2).
#include "kernels.h"

namespace nnic {

// ---------------------------------------------------------------------------------------------------------------
// Dense(512) on the flattened NHWC output of Entropynet.conv3, which the convolution kernel leaves as split fp16 planes
// (x = (hi + lo) / ACT_SCALE).  out[p][o] = bias[o] + sum_f x[p][f] * W[f][o]   (Keras Dense kernel layout [in, out]).
// 64 x 64 output tile per block, 16-deep k chunks through shared memory, 4 x 4 register tile per thread.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DM = 64, DN = 64, DK = 16;

__global__ void __launch_bounds__(256) k_dense_split_in(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo,
                                                        const float* __restrict__ x_f32, int P, int F,
                                                        const float* __restrict__ Wt, const float* __restrict__ bias,
                                                        int NOUT, float* __restrict__ out) {
  __shared__ float As[DK][DM + 4];
  __shared__ float Bs[DK][DN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * DM, n0 = blockIdx.y * DN;
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;      // this thread's 4 x 4 outputs
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;          // A loader: 4 consecutive features of one plane
  const int b_row = tid >> 4, b_col = (tid & 15) * 4;       // B loader: one float4 of W
  float acc[4][4] = {};
  for (int k0 = 0; k0 < F; k0 += DK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + a_row < P) {
      const size_t o = (size_t)(m0 + a_row) * F + k0 + a_k;
      if (x_f32) {
        const float4 v = *reinterpret_cast<const float4*>(x_f32 + o);
        av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
      } else {
        const uint2 h2 = *reinterpret_cast<const uint2*>(x_hi + o), l2 = *reinterpret_cast<const uint2*>(x_lo + o);
        const __half* hh = reinterpret_cast<const __half*>(&h2);
        const __half* ll = reinterpret_cast<const __half*>(&l2);
#pragma unroll
        for (int e = 0; e < 4; ++e) av[e] = join_f32(hh[e], ll[e]);
      }
    }
    const float4 bv = *reinterpret_cast<const float4*>(Wt + (size_t)(k0 + b_row) * NOUT + n0 + b_col);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) As[a_k + e][a_row] = av[e];
    *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      const float a_[4] = {a4.x, a4.y, a4.z, a4.w}, b_[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m < P) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + tn);
      *reinterpret_cast<float4*>(out + (size_t)m * NOUT + n0 + tn) =
          make_float4(__fadd_rn(acc[i][0], bb.x), __fadd_rn(acc[i][1], bb.y), __fadd_rn(acc[i][2], bb.z), __fadd_rn(acc[i][3], bb.w));
    }
  }
}

cudaError_t launch_dense512(const __half* x_hi, const __half* x_lo, const float* x_f32, int P, int F, const float* Wt,
                            const float* bias, float* out, cudaStream_t stream) {
  if (P <= 0 || F <= 0 || F % DK) return cudaErrorInvalidValue;
  k_dense_split_in<<<dim3((P + DM - 1) / DM, 512 / DN), 256, 0, stream>>>(x_hi, x_lo, x_f32, P, F, Wt, bias, 512, out);
  return cudaGetLastError();
}

// Dense(1) + tf.clip_by_value(x, 0, 8): one warp per plane
__global__ void __launch_bounds__(256) k_dense1_clip(const float* __restrict__ hid, int P, const float* __restrict__ w, float b,
                                                     float* __restrict__ out) {
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  float s = 0.f;
  for (int i = lane; i < 512; i += 32) s = fmaf(hid[(size_t)p * 512 + i], w[i], s);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) out[p] = fminf(fmaxf(__fadd_rn(s, b), 0.0f), 8.0f);
}
cudaError_t launch_dense1_clip(const float* hid, int P, const float* w, float b, float* out, cudaStream_t stream) {
  k_dense1_clip<<<(P + 7) / 8, 256, 0, stream>>>(hid, P, w, b, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Noise proxy of quantisation: out = clip(x + u / 255, 0, 1), u ~ U(-0.5, 0.5)  (training.py:87-88).
// `noise` (optional) supplies u explicitly -- the parity tests do, TensorFlow's generator cannot be reproduced -- otherwise
// u comes from Philox4x32-10 with key = seed and counter = element index / 4 (independent of the launch geometry).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

__global__ void __launch_bounds__(256) k_noise_quantise(const float* __restrict__ x, size_t count, unsigned long long seed,
                                                        const float* __restrict__ noise, float* __restrict__ out) {
  const size_t n4 = (count + 3) / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    if (!noise) philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const size_t j = 4 * i + e;
      if (j < count) {
        // u in [-0.5, 0.5): 24 random bits / 2^24 - 0.5
        const float u = noise ? noise[j] : __fsub_rn((float)(r[e] >> 8) * (1.0f / 16777216.0f), 0.5f);
        out[j] = fminf(fmaxf(__fadd_rn(x[j], __fdiv_rn(u, 255.0f)), 0.0f), 1.0f);
      }
    }
  }
}
cudaError_t launch_noise_quantise(const float* x, size_t count, unsigned long long seed, const float* noise, float* out,
                                  int num_sms, cudaStream_t stream) {
  size_t blocks = ((count + 3) / 4 + 255) / 256;
  const size_t cap = (size_t)num_sms * 16;
  k_noise_quantise<<<(unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)), 256, 0, stream>>>(x, count, seed, noise, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// tf.image.ssim for single-channel fp32 images a, b [P][H][W], max_val = 1.
// A block computes a 32 x 32 tile of the (H-10) x (W-10) map: the 42 x 42 inputs go to shared memory, the separable
// Gaussian window (g (x) g equals TensorFlow's 2-D softmax kernel) is applied along rows to the four maps a, b, a*b,
// a*a + b*b, then along columns, and the tile's sum of luminance * contrast-structure goes to partial[block]; a second
// kernel adds the partials of an image in a fixed order (deterministic) and divides by the map size.
// ---------------------------------------------------------------------------------------------------------------
struct SsimWindow { float g[11]; };
constexpr int ST = 32, SI = ST + 10;

__global__ void __launch_bounds__(256) k_ssim_tiles(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                    int tiles_x, int tiles_y, const __grid_constant__ SsimWindow win,
                                                    float* __restrict__ partial) {
  __shared__ float sa[SI][SI + 1], sb[SI][SI + 1];
  __shared__ float hm[4][SI][ST + 1];
  __shared__ float red[8];
  const int tile = blockIdx.x % (tiles_x * tiles_y), p = blockIdx.x / (tiles_x * tiles_y);
  const int y0 = (tile / tiles_x) * ST, x0 = (tile % tiles_x) * ST;
  const int Ho = H - 10, Wo = W - 10;
  const float* pa = a + (size_t)p * H * W;
  const float* pb = b + (size_t)p * H * W;
  for (int i = threadIdx.x; i < SI * SI; i += 256) {
    const int r = i / SI, c = i - r * SI;
    const int y = y0 + r, x = x0 + c;
    const bool ok = y < H && x < W;
    sa[r][c] = ok ? pa[(size_t)y * W + x] : 0.f;
    sb[r][c] = ok ? pb[(size_t)y * W + x] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SI * ST; i += 256) {
    const int r = i / ST, c = i - r * ST;
    float m0 = 0.f, m1 = 0.f, mxy = 0.f, msq = 0.f;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float va = sa[r][c + t], vb = sb[r][c + t], g = win.g[t];
      m0 = fmaf(g, va, m0); m1 = fmaf(g, vb, m1);
      mxy = fmaf(g, __fmul_rn(va, vb), mxy);
      msq = fmaf(g, __fadd_rn(__fmul_rn(va, va), __fmul_rn(vb, vb)), msq);
    }
    hm[0][r][c] = m0; hm[1][r][c] = m1; hm[2][r][c] = mxy; hm[3][r][c] = msq;
  }
  __syncthreads();
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  float sum = 0.f;
  for (int i = threadIdx.x; i < ST * ST; i += 256) {
    const int r = i / ST, c = i - r * ST;
    if (y0 + r < Ho && x0 + c < Wo) {
      float m0 = 0.f, m1 = 0.f, mxy = 0.f, msq = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) {
        const float g = win.g[t];
        m0 = fmaf(g, hm[0][r + t][c], m0); m1 = fmaf(g, hm[1][r + t][c], m1);
        mxy = fmaf(g, hm[2][r + t][c], mxy); msq = fmaf(g, hm[3][r + t][c], msq);
      }
      const float num0 = __fmul_rn(__fmul_rn(m0, m1), 2.0f), den0 = __fadd_rn(__fmul_rn(m0, m0), __fmul_rn(m1, m1));
      const float lum = __fdiv_rn(__fadd_rn(num0, c1), __fadd_rn(den0, c1));
      const float num1 = __fmul_rn(mxy, 2.0f);
      const float cs = __fdiv_rn(__fadd_rn(__fsub_rn(num1, num0), c2), __fadd_rn(__fsub_rn(msq, den0), c2));
      sum += __fmul_rn(lum, cs);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void k_ssim_finish(const float* __restrict__ partial, int tiles_per_image, int P, double inv_count, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double s = 0.0;
  for (int t = 0; t < tiles_per_image; ++t) s += (double)partial[(size_t)p * tiles_per_image + t];
  out[p] = (float)(s * inv_count);
}

cudaError_t launch_ssim(const float* a, const float* b, int P, int H, int W, float* partial, float* out, cudaStream_t stream) {
  if (H < 11 || W < 11 || P <= 0) return cudaErrorInvalidValue;
  SsimWindow win;
  double g[11], tot = 0.0;
  for (int i = 0; i < 11; ++i) { const double d = i - 5.0; g[i] = exp(-0.5 * d * d / (1.5 * 1.5)); tot += g[i]; }
  for (int i = 0; i < 11; ++i) win.g[i] = (float)(g[i] / tot);
  const int Ho = H - 10, Wo = W - 10;
  const int tiles_x = (Wo + ST - 1) / ST, tiles_y = (Ho + ST - 1) / ST;
  const long long blocks = (long long)tiles_x * tiles_y * P;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  k_ssim_tiles<<<(unsigned)blocks, 256, 0, stream>>>(a, b, H, W, tiles_x, tiles_y, win, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_ssim_finish<<<(P + 127) / 128, 128, 0, stream>>>(partial, tiles_x * tiles_y, P, 1.0 / ((double)Ho * Wo), out);
  return cudaGetLastError();
}
size_t ssim_partial_count(int P, int H, int W) {
  const int tiles_x = (W - 10 + ST - 1) / ST, tiles_y = (H - 10 + ST - 1) / ST;
  return (size_t)tiles_x * tiles_y * (size_t)P;
}

}  // namespace nnic
