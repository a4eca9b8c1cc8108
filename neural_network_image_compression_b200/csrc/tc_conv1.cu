// colour transform + conv1 on the tensor cores.
//
// Reference: Encoder.__call__ lines 39-41 (x/255, convert_to_colourspace, utils.py:64-77) and BaseEncoder.conv1
// (encoder.py:10,20): Conv2D(32, 5, 2, 'SAME', leaky_relu) on ONE input channel.
//
// As a GEMM: D[128 pixels, 32 channels] = A[128 pixels, K = 25 taps (+7 zero columns)] x W[K, 32].  K is tiny, so
// the layer is bound by building A and by the 96 B/pixel it writes, not by the MMAs.  A has no TMA-loadable form
// (one input channel), so eight "builder" warps (four groups on different tiles) construct it: they load the 35 x 19 input patch of a 16 x 8 output
// tile (colour transform applied on the fly, exactly as the reference orders its fp32 ops), and every pixel's 25 taps
// are split into fp16 hi/lo and written straight into the SWIZZLE_64B K-major layout the tensor core reads
// (64-byte rows).  The group's first warp then issues 2 k-steps x 2 MMAs per tile (A_hi x [W_hi|W_lo], A_lo x W_hi) into a TMEM
// slot [main 32 | corr 32]; eight epilogue warps (TMEM lane group x channel half) add bias, apply leaky_relu, split
// to fp16 hi/lo and write each pixel's 16 channels with one 256-bit store per plane.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kBuilderWarps = 8, kEpiWarps = 8;
constexpr int kEpiWarp0 = kBuilderWarps;              // warps 0-7 builders (the first warp of a group also issues its MMAs), 8-15 epilogue
constexpr int kThreads = (kBuilderWarps + kEpiWarps) * 32;       // 512: four warps per scheduler, 128 registers per thread
constexpr int CO = 32;                                // output channels
constexpr int PH = 2 * kTileRows + 3, PW = 2 * kTileCols + 3;    // 35 x 19 input pixels
constexpr int PPITCH = 12;                            // words per row of one column-parity plane of the patch (10 / 9 used):
                                                      // two patch rows = 24 words, so the 4 x 8 output pixels a warp builds
                                                      // read 32 distinct banks
constexpr int PPLANE = PH * PPITCH;                   // words per parity plane
constexpr int A_TILE = kTileM * 64;                   // 8 KB: 128 rows x 32 fp16
constexpr int STAGE_BYTES = 2 * A_TILE;               // hi + lo
constexpr int GROUPS = 4;                             // builder groups (2 warps each) working on different tiles
constexpr int GTHREADS = kBuilderWarps * 32 / GROUPS;  // 64
constexpr int STAGES = 2 * GROUPS;
constexpr int PATCH_BYTES = (2 * PPLANE * 4 + 15) / 16 * 16;
constexpr int W_TILE = CO * 64;                       // 2 KB: 32 rows x 32 fp16
constexpr int W_SET = 2 * W_TILE;                     // [W_hi | W_lo]
constexpr int SLOT_COLS = 2 * CO, SLOTS = 8, TMEM_COLS = 512;
constexpr int PATCH_OFF = STAGES * STAGE_BYTES;
constexpr int W_OFF = (PATCH_OFF + GROUPS * PATCH_BYTES + 1023) / 1024 * 1024;   // swizzle patterns are functions of the absolute address
constexpr int BAR_OFF = (W_OFF + 2 * W_SET + 1023) / 1024 * 1024;
constexpr int SMEM_BYTES = BAR_OFF + 512 + 2 * CO * 4 + 1024;

// 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row m lives at chunk j ^ ((m >> 1) & 3)
__device__ __forceinline__ uint32_t sw64(int m, int j) { return (uint32_t)(m * 64 + ((j ^ ((m >> 1) & 3)) << 4)); }

// x.astype(float32)/255 for a byte x, bit-identical to the IEEE division: q = RN(x * RN(1/255)) refined by one
// residual step (r = fma(-q, 255, x) is exact, RN(q + r/255) is the correctly rounded quotient; checked for all 256
// inputs by tests/test_host.py::test_div255_refinement and, on the device, by the colour-plane parity tests).
__device__ __forceinline__ float div255(uint32_t byte) {
  const float x = (float)byte;
  const float y = 0x1.010102p-8f;                       // RN(1/255)
  const float q = __fmul_rn(x, y);
  return fmaf(fmaf(-q, 255.0f, x), y, q);
}

// Persistent-loop tile coordinates (plane, tile row, tile column) advanced without divisions.
struct TileIter {
  int p, ty, tx;       // current tile
  int sp, sy, sx;      // stride decomposed the same way
  __device__ __forceinline__ void init(int t0, int stride, int tiles_x, int tiles_y) {
    const int tpp = tiles_x * tiles_y;
    p = t0 / tpp; int r = t0 - p * tpp; ty = r / tiles_x; tx = r - ty * tiles_x;
    sp = stride / tpp; r = stride - sp * tpp; sy = r / tiles_x; sx = r - sy * tiles_x;
  }
  __device__ __forceinline__ void next(int tiles_x, int tiles_y) {
    tx += sx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    ty += sy; if (ty >= tiles_y) { ty -= tiles_y; ++p; }
    p += sp;
  }
};

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GTHREADS) : "memory"); }

template <int IN_KIND /*0 rgb u8, 1 f32 planes*/>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv1(const __grid_constant__ TcConv1Params prm, int tiles_x, int tiles_y, int num_tiles, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  const WaitCtx wc{error_flag, prm.wait_timeout};
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* w_base = smem + W_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty_bar = bars;                   // [STAGES]  MMAs done reading a stage -> builders
  uint64_t* slot_full = bars + STAGES;          // [SLOTS]   tile k uses stage k % STAGES and TMEM slot k % SLOTS
  uint64_t* slot_empty = slot_full + SLOTS;     // [SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + SLOTS);
  float* bias_s = reinterpret_cast<float*>(smem + BAR_OFF + 512);      // [2][32]
  static_assert((STAGES + 2 * SLOTS) * 8 + 4 <= 512, "barrier area too small");
  static_assert(STAGES == SLOTS, "a tile's stage and TMEM slot share index and phase");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = prm.N, H = prm.H, W = prm.W, Ho = prm.Ho, Wo = prm.Wo;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * CO; i += kThreads) bias_s[i] = prm.bias[i] * ACT_SCALE;   // the epilogue works on scaled values
  // weights: [set][W_hi | W_lo], each [32 channels][32 taps] fp16 K-major, SWIZZLE_64B; 16-byte chunks copied by all threads
  for (int i = threadIdx.x; i < 2 * 2 * CO * 4; i += kThreads) {
    const int j = i & 3, row = (i >> 2) & (CO - 1), part = (i >> 7) & 1, set = i >> 8;
    const uint4 v = *reinterpret_cast<const uint4*>((part ? prm.w_lo : prm.w_hi) + ((size_t)(set * CO + row) * 32 + j * 8));
    *reinterpret_cast<uint4*>(w_base + set * W_SET + part * W_TILE + sw64(row, j)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_plane = tiles_x * tiles_y;

  if (warp < kBuilderWarps) {
    // ===================== builders: input patch -> im2col tile in the MMA layout, then its MMAs =====================
    // Four groups of two warps work on different tiles (tile counter % GROUPS), so the global-load latency of one
    // group's patch overlaps the im2col arithmetic of the others.  Tile counter k uses smem stage k % STAGES.
    const int g = warp / (kBuilderWarps / GROUPS);
    const int gt = threadIdx.x - g * GTHREADS;            // 0..63 inside the group
    // packed fp16 (hi | lo << 16) of every patch pixel, even columns in plane 0 and odd columns in plane 1
    uint32_t* patch = reinterpret_cast<uint32_t*>(smem + PATCH_OFF + g * PATCH_BYTES);
    // the patch pixels this thread converts: i = gt + 64 q -> (row, column); fixed for the whole kernel
    constexpr int PER = (PH * PW + GTHREADS - 1) / GTHREADS;     // 11
    int rel[PER];                                          // byte (rgb) or element (planes) offset inside the image, relative to the patch origin
    int16_t prc[PER];                                      // row | column << 8
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = gt + q * GTHREADS;
      const int pr = i / PW, pc = i - pr * PW;
      rel[q] = (pr * W + pc) * (IN_KIND == 0 ? 3 : 1);
      prc[q] = (int16_t)(pr | (pc << 8));
    }
    TileIter it;
    it.init(blockIdx.x + g * gridDim.x, GROUPS * gridDim.x, tiles_x, tiles_y);
    const int P = 3 * N;
    // Raw pixels of a tile: unconditional loads from an always-valid address keep all of them in flight together;
    // out-of-image pixels are zeroed at conversion (= TF SAME padding).  The loads of tile k+1 are issued before tile
    // k is built, so their latency hides behind the im2col work.
    uint32_t c0[PER], c1[PER], c2[PER];
    float fval[PER];
    uint32_t okmask = 0;
    auto load_raw = [&](const TileIter& t) {
      const int plane = t.p / N, n = t.p - plane * N;
      const int iy0 = t.ty * kTileRows * 2 - prm.pad_t, ix0 = t.tx * kTileCols * 2 - prm.pad_l;
      const bool interior = iy0 >= 0 && iy0 + PH <= H && ix0 >= 0 && ix0 + PW <= W;
      const uint8_t* img_u8 = prm.rgb + ((size_t)n * H * W + (ptrdiff_t)iy0 * W + ix0) * 3;
      const float* img_f = prm.planes + ((size_t)t.p * H * W + (ptrdiff_t)iy0 * W + ix0);
      okmask = 0;
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int pr = prc[q] & 0xff, pc = prc[q] >> 8;
        const bool in_patch = q < PER - 1 || gt + q * GTHREADS < PH * PW;
        const bool ok = in_patch && (interior || ((unsigned)(iy0 + pr) < (unsigned)H && (unsigned)(ix0 + pc) < (unsigned)W));
        okmask |= (ok ? 1u : 0u) << q;
        if (IN_KIND == 0) {
          const uint8_t* px = ok ? img_u8 + rel[q] : prm.rgb;
          c0[q] = px[0]; c1[q] = px[1]; c2[q] = px[2];
        } else {
          const float* pf = ok ? img_f + rel[q] : prm.planes;
          fval[q] = *pf;
        }
      }
    };
    if (it.p < P) load_raw(it);
    int counter = g;
    for (; it.p < P; counter += GROUPS) {
      const int stage = counter % STAGES;
      const uint32_t phase = (counter / STAGES) & 1;
      const int p = it.p;
      const int plane = p / N;
      const float k0 = prm.cc.k[plane][0], k1 = prm.cc.k[plane][1], k2 = prm.cc.k[plane][2], off = prm.cc.off[plane];
      group_barrier(g);                          // the previous tile of this group no longer reads the patch
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        if (q < PER - 1 || gt + q * GTHREADS < PH * PW) {
          const int pr = prc[q] & 0xff, pc = prc[q] >> 8;
          const bool ok = (okmask >> q) & 1u;
          float v;
          if (IN_KIND == 0) {
            // x.astype(float32)/255, then (t0*k0 + t1*k1) + t2*k2 with separate roundings, then + offset
            const float r_ = div255(c0[q]), g_ = div255(c1[q]), b_ = div255(c2[q]);
            v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r_, k0), __fmul_rn(g_, k1)), __fmul_rn(b_, k2)), off);
          } else {
            v = fval[q];
          }
          v = ok ? v : 0.0f;
          __half hi, lo;
          split_f32(v, hi, lo);                 // split once per input pixel; every tap that uses it copies the halves
          patch[(pc & 1) * PPLANE + pr * PPITCH + (pc >> 1)] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
      }
      it.next(tiles_x, tiles_y);
      if (it.p < P) load_raw(it);                // next tile of this group
      mbar_wait(&empty_bar[stage], phase ^ 1, wc, 1);      // the MMAs that read this stage are done
      group_barrier(g);
      // K column k = kh*5 + kw for k < 25, zero above; each thread builds the rows of two pixels
      uint8_t* a_hi = stage_base + stage * STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_TILE;
#pragma unroll
      for (int mm = 0; mm < 2; ++mm) {
        const int m = gt + mm * GTHREADS;
        const int r = m >> 3, c = m & 7;
        const uint32_t* prow = &patch[(2 * r) * PPITCH + c];       // tap (kh, kw): plane kw & 1, row + kh, column + kw / 2
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ka = j * 8 + 2 * e, kb = ka + 1;
            const uint32_t a = ka < 25 ? prow[((ka % 5) & 1) * PPLANE + (ka / 5) * PPITCH + ((ka % 5) >> 1)] : 0u;
            const uint32_t b = kb < 25 ? prow[((kb % 5) & 1) * PPLANE + (kb / 5) * PPITCH + ((kb % 5) >> 1)] : 0u;
            hw[e] = __byte_perm(a, b, 0x5410);   // hi halves of taps ka, kb
            lw[e] = __byte_perm(a, b, 0x7632);   // lo halves
          }
          *reinterpret_cast<uint4*>(a_hi + sw64(m, j)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(a_lo + sw64(m, j)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to tcgen05.mma
      group_barrier(g);                          // both warps of the group have written their rows
      if ((warp & (kBuilderWarps / GROUPS - 1)) == 0) {
        // ---- MMA issue: A_hi x [W_hi | W_lo] and A_lo x W_hi, two k-steps, into TMEM slot `stage`
        mbar_wait(&slot_empty[stage], phase ^ 1, wc, 2);
        tc_fence_after();
        if (elect_one()) {
          const int set = p < N ? 0 : 1;
          const uint32_t d_tmem = tmem_base + stage * SLOT_COLS;
          const uint64_t da_hi = make_smem_desc<64>(smem_u32(a_hi));
          const uint64_t da_lo = da_hi + (uint64_t)(A_TILE >> 4);
          const uint64_t w_hl = make_smem_desc<64>(smem_u32(w_base) + set * W_SET);          // [W_hi | W_lo]: 64 rows
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            umma_f16(d_tmem, da_hi + 2 * ks, w_hl + 2 * ks, make_idesc(2 * CO), ks ? 1u : 0u);
            umma_f16(d_tmem + CO, da_lo + 2 * ks, w_hl + 2 * ks, make_idesc(CO), 1u);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&slot_full[stage]);
        }
        __syncwarp();
      }
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
    // and barrier latencies overlap the bias / leaky / split / store work instead of adding to it.  The loop is
    // unrolled by two so that the two register buffers swap roles without copies.
    int slot = 0; uint32_t slot_phase = 0;
    uint32_t am[HALF], ac[HALF], bm[HALF], bc[HALF];
    auto issue_loads = [&](uint32_t* vm, uint32_t* vc) {
      mbar_wait(&slot_full[slot], slot_phase, wc, 4);
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
    const int m = lg * 32 + lane;
    const int my = m >> 3, mx = m & 7;
    TileIter it;
    it.init(blockIdx.x, gridDim.x, tiles_x, tiles_y);
    // one tile: values are kept multiplied by ACT_SCALE (a power of two: bias add, leaky_relu and the hi/lo split
    // commute with it bit for bit), which saves the scaling multiplies of the split
    auto finish_tile = [&](const uint32_t* vm, const uint32_t* vc) {
      const int p = it.p;
      const int set = p < N ? 0 : 1;
      const float inv16 = prm.inv_scale[set] * ACT_SCALE;
      const float4* bs4 = reinterpret_cast<const float4*>(bias_s + set * CO + ch0);
      uint32_t h[HALF / 2], l[HALF / 2];
#pragma unroll
      for (int i = 0; i < HALF; i += 4) {
        const float4 b = bs4[i / 4];
        float v0 = fmaf(__fadd_rn(__uint_as_float(vm[i]), __uint_as_float(vc[i])), inv16, b.x);
        float v1 = fmaf(__fadd_rn(__uint_as_float(vm[i + 1]), __uint_as_float(vc[i + 1])), inv16, b.y);
        float v2 = fmaf(__fadd_rn(__uint_as_float(vm[i + 2]), __uint_as_float(vc[i + 2])), inv16, b.z);
        float v3 = fmaf(__fadd_rn(__uint_as_float(vm[i + 3]), __uint_as_float(vc[i + 3])), inv16, b.w);
        v0 = fmaxf(v0, __fmul_rn(v0, LEAKY_ALPHA));
        v1 = fmaxf(v1, __fmul_rn(v1, LEAKY_ALPHA));
        v2 = fmaxf(v2, __fmul_rn(v2, LEAKY_ALPHA));
        v3 = fmaxf(v3, __fmul_rn(v3, LEAKY_ALPHA));
        split2_scaled(v0, v1, h[i / 2], l[i / 2]);
        split2_scaled(v2, v3, h[i / 2 + 1], l[i / 2 + 1]);
      }
      // 16 channels = 32 bytes per fp16 plane: one 256-bit store each (a full sector per thread)
      const int y = it.ty * kTileRows + my, x = it.tx * kTileCols + mx;
      if (y < Ho && x < Wo) {
        const size_t o = (((size_t)p * prm.Hs + y) * prm.Ws + x) * CO + ch0;
        st_global_v8(prm.out_hi + o, h);
        st_global_v8(prm.out_lo + o, l);
      }
    };
    const int P = 3 * N;
    if (it.p < P) { issue_loads(am, ac); release_slot(); }
    while (it.p < P) {
      TileIter nx = it; nx.next(tiles_x, tiles_y);
      if (nx.p < P) issue_loads(bm, bc);
      finish_tile(am, ac);
      it = nx;
      if (it.p >= P) break;
      release_slot();
      nx.next(tiles_x, tiles_y);
      if (nx.p < P) issue_loads(am, ac);
      finish_tile(bm, bc);
      it = nx;
      if (it.p < P) release_slot();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace

cudaError_t launch_tc_conv1(const TcConv1Params& prm, int num_sms, int* error_flag, cudaStream_t stream) {
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_conv1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
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
