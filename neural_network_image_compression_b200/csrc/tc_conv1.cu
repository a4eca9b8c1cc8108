// colour transform + conv1 on the tensor cores.
//
// Reference: Encoder.__call__ lines 39-41 (x/255, convert_to_colourspace, utils.py:64-77) and BaseEncoder.conv1
// (encoder.py:10,20): Conv2D(32, 5, 2, 'SAME', leaky_relu) on ONE input channel.
//
// As a GEMM: D[128 pixels, 32 channels] = A[128 pixels, K = 25 taps (+7 zero columns)] x W[K, 32].  K is tiny, so
// the layer is bound by building A and by the 96 B/pixel it writes, not by the MMAs.  A has no TMA-loadable form
// (one input channel), so eight "builder" warps construct it: they load the 35 x 19 input patch of a 16 x 8 output
// tile (colour transform applied on the fly, exactly as the reference orders its fp32 ops), and every pixel's 25 taps
// are split into fp16 hi/lo and written straight into the SWIZZLE_64B K-major layout the tensor core reads
// (64-byte rows).  One MMA warp issues 2 k-steps x 2 MMAs per tile (A_hi x [W_hi|W_lo], A_lo x W_hi) into a TMEM
// slot [main 32 | corr 32]; four epilogue warps add bias, apply leaky_relu, split to fp16 hi/lo and store through
// a per-warp staging buffer with fully coalesced 16-byte stores.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kBuilderWarps = 8, kEpiWarps = 8;
constexpr int kMmaWarp = kBuilderWarps;               // warps 0-7 builders, 8 MMA, 9-12 epilogue
constexpr int kEpiWarp0 = kBuilderWarps + 1;
constexpr int kThreads = (kBuilderWarps + 1 + kEpiWarps) * 32;   // 544
constexpr int CO = 32;                                // output channels
constexpr int PH = 2 * kTileRows + 3, PW = 2 * kTileCols + 3;    // 35 x 19 input pixels
constexpr int PPITCH = PW + 1;
constexpr int A_TILE = kTileM * 64;                   // 8 KB: 128 rows x 32 fp16
constexpr int STAGE_BYTES = 2 * A_TILE;               // hi + lo
constexpr int GROUPS = 4;                             // builder groups (2 warps each) working on different tiles
constexpr int GTHREADS = kBuilderWarps * 32 / GROUPS;  // 64
constexpr int STAGES = 2 * GROUPS;
constexpr int PATCH_BYTES = (PH * PPITCH * 4 + 15) / 16 * 16;
constexpr int W_TILE = CO * 64;                       // 2 KB: 32 rows x 32 fp16
constexpr int W_SET = 2 * W_TILE;                     // [W_hi | W_lo]
constexpr int SLOT_COLS = 2 * CO, SLOTS = 8, TMEM_COLS = 512;
constexpr int PATCH_OFF = STAGES * STAGE_BYTES;
constexpr int W_OFF = (PATCH_OFF + GROUPS * PATCH_BYTES + 1023) / 1024 * 1024;   // swizzle patterns are functions of the absolute address
constexpr int STG_OFF = (W_OFF + 2 * W_SET + 1023) / 1024 * 1024;
constexpr int STG_WARP = 32 * 32;                     // 32 pixels x 16 fp16
constexpr int LUT_BYTES = 256 * 4;                    // float(x)/255 for every byte value
constexpr int BAR_OFF = STG_OFF + kEpiWarps * STG_WARP;
constexpr int SMEM_BYTES = BAR_OFF + 512 + 2 * CO * 4 + LUT_BYTES + 1024;

// 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row m lives at chunk j ^ ((m >> 1) & 3)
__device__ __forceinline__ uint32_t sw64(int m, int j) { return (uint32_t)(m * 64 + ((j ^ ((m >> 1) & 3)) << 4)); }

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GTHREADS) : "memory"); }

template <int IN_KIND /*0 rgb u8, 1 f32 planes*/>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv1(const __grid_constant__ TcConv1Params prm, int tiles_x, int tiles_y, int num_tiles, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* w_base = smem + W_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_bar = bars;                    // [STAGES]  builders -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> builders
  uint64_t* slot_full = bars + 2 * STAGES;      // [SLOTS]
  uint64_t* slot_empty = slot_full + SLOTS;     // [SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + SLOTS);
  float* bias_s = reinterpret_cast<float*>(smem + BAR_OFF + 512);      // [2][32]
  float* div255 = bias_s + 2 * CO;                                     // [256]: x.astype(float32)/255 as an IEEE division
  static_assert((2 * STAGES + 2 * SLOTS) * 8 + 4 <= 512, "barrier area too small");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = prm.N, H = prm.H, W = prm.W, Ho = prm.Ho, Wo = prm.Wo;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kBuilderWarps / GROUPS); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * CO; i += kThreads) bias_s[i] = prm.bias[i];
  for (int i = threadIdx.x; i < 256; i += kThreads) div255[i] = __fdiv_rn((float)i, 255.0f);
  // weights: [set][W_hi | W_lo], each [32 channels][32 taps] fp16 K-major, SWIZZLE_64B; 16-byte chunks copied by all threads
  for (int i = threadIdx.x; i < 2 * 2 * CO * 4; i += kThreads) {
    const int j = i & 3, row = (i >> 2) & (CO - 1), part = (i >> 7) & 1, set = i >> 8;
    const uint4 v = *reinterpret_cast<const uint4*>((part ? prm.w_lo : prm.w_hi) + ((size_t)(set * CO + row) * 32 + j * 8));
    *reinterpret_cast<uint4*>(w_base + set * W_SET + part * W_TILE + sw64(row, j)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_plane = tiles_x * tiles_y;

  if (warp < kBuilderWarps) {
    // ===================== builders: input patch -> im2col tile in the MMA layout =====================
    // Four groups of two warps work on different tiles (tile counter % GROUPS), so the global-load latency of one
    // group's patch overlaps the im2col arithmetic of the others.  Tile counter k uses smem stage k % STAGES.
    const int g = warp / (kBuilderWarps / GROUPS);
    const int gt = threadIdx.x - g * GTHREADS;            // 0..63 inside the group
    uint32_t* patch = reinterpret_cast<uint32_t*>(smem + PATCH_OFF + g * PATCH_BYTES);   // fp16 hi | lo << 16 of every patch pixel
    int counter = g;
    for (int t = blockIdx.x + g * gridDim.x; t < num_tiles; t += GROUPS * gridDim.x, counter += GROUPS) {
      const int stage = counter % STAGES;
      const uint32_t phase = (counter / STAGES) & 1;
      const int txy = t % tiles_per_plane, p = t / tiles_per_plane;
      const int plane = p / N, n = p - plane * N;
      const int iy0 = (txy / tiles_x) * kTileRows * 2 - prm.pad_t, ix0 = (txy % tiles_x) * kTileCols * 2 - prm.pad_l;
      const float k0 = prm.cc.k[plane][0], k1 = prm.cc.k[plane][1], k2 = prm.cc.k[plane][2], off = prm.cc.off[plane];
      group_barrier(g);                          // the previous tile of this group no longer reads the patch
      // raw loads first (all in flight together), conversion afterwards
      constexpr int PER = (PH * PW + GTHREADS - 1) / GTHREADS;     // 11
      uint32_t raw[PER];
      float fval[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = gt + q * GTHREADS;
        const int pr = i / PW, pc = i - pr * PW;
        const int iy = iy0 + pr, ix = ix0 + pc;
        raw[q] = 0xffffffffu; fval[q] = 0.0f;
        if (i < PH * PW && iy >= 0 && iy < H && ix >= 0 && ix < W) {
          if (IN_KIND == 0) {
            const uint8_t* px = prm.rgb + (((size_t)n * H + iy) * W + ix) * 3;
            raw[q] = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
          } else {
            fval[q] = prm.planes[((size_t)p * H + iy) * W + ix];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = gt + q * GTHREADS;
        if (i < PH * PW) {
          const int pr = i / PW, pc = i - pr * PW;
          float v = fval[q];
          if (IN_KIND == 0 && raw[q] != 0xffffffffu) {
            // x.astype(float32)/255 (table of IEEE quotients), then (t0*k0 + t1*k1) + t2*k2 with separate roundings, then + offset
            const float r_ = div255[raw[q] & 0xff], g_ = div255[(raw[q] >> 8) & 0xff], b_ = div255[(raw[q] >> 16) & 0xff];
            v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r_, k0), __fmul_rn(g_, k1)), __fmul_rn(b_, k2)), off);
          }
          __half hi, lo;
          split_f32(v, hi, lo);                 // split once per input pixel; every tap that uses it copies the halves
          patch[pr * PPITCH + pc] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
      }
      mbar_wait(&empty_bar[stage], phase ^ 1, error_flag, 1);      // the MMAs that read this stage are done
      group_barrier(g);
      // K column k = kh*5 + kw for k < 25, zero above; each thread builds the rows of two pixels
      uint8_t* a_hi = stage_base + stage * STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_TILE;
#pragma unroll
      for (int mm = 0; mm < 2; ++mm) {
        const int m = gt + mm * GTHREADS;
        const int r = m >> 3, c = m & 7;
        const uint32_t* prow = &patch[(2 * r) * PPITCH + 2 * c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ka = j * 8 + 2 * e, kb = ka + 1;
            const uint32_t a = ka < 25 ? prow[(ka / 5) * PPITCH + (ka % 5)] : 0u;
            const uint32_t b = kb < 25 ? prow[(kb / 5) * PPITCH + (kb % 5)] : 0u;
            hw[e] = __byte_perm(a, b, 0x5410);   // hi halves of taps ka, kb
            lw[e] = __byte_perm(a, b, 0x7632);   // lo halves
          }
          *reinterpret_cast<uint4*>(a_hi + sw64(m, j)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(a_lo + sw64(m, j)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_wide = make_idesc(2 * CO);
    constexpr uint32_t idesc_narrow = make_idesc(CO);
    const uint32_t stage_u32 = smem_u32(stage_base), w_u32 = smem_u32(w_base);
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int p = t / tiles_per_plane;
      const int set = p < N ? 0 : 1;
      mbar_wait(&slot_empty[slot], slot_phase ^ 1, error_flag, 2);
      mbar_wait(&full_bar[stage], phase, error_flag, 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + slot * SLOT_COLS;
      const uint64_t a_hi = make_smem_desc<64>(stage_u32 + stage * STAGE_BYTES);
      const uint64_t a_lo = a_hi + (uint64_t)(A_TILE >> 4);
      const uint64_t w_hl = make_smem_desc<64>(w_u32 + set * W_SET);          // [W_hi | W_lo]: 64 rows
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, ks ? 1u : 0u);
          umma_f16(d_tmem + CO, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&slot_full[slot]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
      if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    // 8 warps: warp w reads TMEM lanes 32*(w%4)..+31 and 16 of the 32 channels
    constexpr int HALF = CO / 2;
    const int ew = warp - kEpiWarp0;
    const int lg = warp & 3;                  // TMEM lane group
    const int hf = ew >> 2;
    const int ch0 = hf * HALF;
    // Software-pipelined: the TMEM loads of tile t+1 are issued before the arithmetic of tile t, so the TMEM
    // and barrier latencies overlap the bias / leaky / split / store work instead of adding to it.
    int slot = 0; uint32_t slot_phase = 0;
    uint32_t cm[HALF], cc_[HALF], nm[HALF], nc[HALF];
    auto issue_loads = [&](uint32_t* vm, uint32_t* vc) {
      mbar_wait(&slot_full[slot], slot_phase, error_flag, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * SLOT_COLS + ch0;
      tmem_ld16_nowait(taddr, vm);
      tmem_ld16_nowait(taddr + CO, vc);
    };
    auto release_slot = [&]() {
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[slot]);
      if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
    };
    if ((int)blockIdx.x < num_tiles) { issue_loads(cm, cc_); release_slot(); }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int txy = t % tiles_per_plane, p = t / tiles_per_plane;
      const int set = p < N ? 0 : 1;
      const int tY0 = (txy / tiles_x) * kTileRows, tX0 = (txy % tiles_x) * kTileCols;
      const bool has_next = t + (int)gridDim.x < num_tiles;
      if (has_next) issue_loads(nm, nc);
      const float inv_scale = prm.inv_scale[set];
      const float* bs = bias_s + set * CO + ch0;
      uint32_t h[HALF / 2], l[HALF / 2];
#pragma unroll
      for (int i = 0; i < HALF; i += 2) {
        float v0 = fmaf(__fadd_rn(__uint_as_float(cm[i]), __uint_as_float(cc_[i])), inv_scale, bs[i]);
        float v1 = fmaf(__fadd_rn(__uint_as_float(cm[i + 1]), __uint_as_float(cc_[i + 1])), inv_scale, bs[i + 1]);
        v0 = fmaxf(v0, __fmul_rn(v0, LEAKY_ALPHA));
        v1 = fmaxf(v1, __fmul_rn(v1, LEAKY_ALPHA));
        split2_f32(v0, v1, h[i / 2], l[i / 2]);
      }
      // 16 channels = 32 bytes per fp16 plane: one 256-bit store each (a full sector per thread)
      const int m = lg * 32 + lane;
      const int y = tY0 + (m >> 3), x = tX0 + (m & 7);
      if (y < Ho && x < Wo) {
        const size_t o = (((size_t)p * Ho + y) * Wo + x) * CO + ch0;
        st_global_v8(prm.out_hi + o, h);
        st_global_v8(prm.out_lo + o, l);
      }
      if (has_next) {
        release_slot();
#pragma unroll
        for (int i = 0; i < HALF; ++i) { cm[i] = nm[i]; cc_[i] = nc[i]; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace

cudaError_t launch_tc_conv1(const TcConv1Params& prm, int num_sms, int* error_flag, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_conv1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles_x = (prm.Wo + kTileCols - 1) / kTileCols, tiles_y = (prm.Ho + kTileRows - 1) / kTileRows;
  const long long tiles = (long long)tiles_x * tiles_y * 3 * prm.N;
  if (tiles <= 0 || tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int grid = tiles < num_sms ? (int)tiles : num_sms;
  if (prm.rgb) k_tc_conv1<0><<<grid, kThreads, SMEM_BYTES, stream>>>(prm, tiles_x, tiles_y, (int)tiles, error_flag);
  else k_tc_conv1<1><<<grid, kThreads, SMEM_BYTES, stream>>>(prm, tiles_x, tiles_y, (int)tiles, error_flag);
  return cudaGetLastError();
}

}  // namespace nnic
