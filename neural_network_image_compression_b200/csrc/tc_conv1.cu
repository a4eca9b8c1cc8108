// colour transform + conv1 on the tensor cores.
//
// Reference: Encoder.__call__ lines 39-41 (x/255, convert_to_colourspace, utils.py:64-77) and BaseEncoder.conv1
// (encoder.py:10,20): Conv2D(32, 5, 2, 'SAME', leaky_relu) on ONE input channel.
//
// As a GEMM: D[128 pixels, 32 channels] = A[128 pixels, K = 25 taps] x W[K, 32].  K is tiny, so the layer is bound by
// building A and by the 96 B/pixel it writes, not by the MMAs.  A has no TMA-loadable form (one input channel), so
// "builder" warps construct it.  Round 1 looped over (plane, tile): every RGB pixel was fetched three times with byte
// loads, /255 was evaluated nine times per pixel, and each im2col row cost ~100 instructions (25 word loads + 25 byte
// permutes to separate the packed hi|lo halves); ncu showed the builders issuing one instruction per 9.5 cycles per warp
// and the kernel at 38 % of the copy bandwidth (profiles/r1_ncu_final_all_kernels.txt).  This version:
//   * one work item = one 16 x 8 output tile of one IMAGE, all three colour planes: the 35 x 19 RGB patch is fetched once
//     with aligned 32-bit loads (9 per thread instead of 99 byte loads), staged in shared memory, and every pixel is
//     normalised once and projected onto the three planes;
//   * the converted patch is kept as separate fp16 hi and lo planes, and the K axis is laid out as 5 kernel rows x 6 slots
//     (5 taps + one zero-weight slot): the five taps of a kernel row are five CONSECUTIVE input pixels starting at an even
//     column, so a row of A is 15 aligned 32-bit shared loads per half and no permutes (the weight matrix carries zeros in
//     slots 5, 11, 17, 23, 29, 30, 31);
//   * four builder groups (two warps each) work on different items; a group's first warp issues the MMAs of its three
//     planes (2 k-steps x [A_hi x [W_hi|W_lo], A_lo x W_hi]) into TMEM slots [main 32 | corr 32];
//   * eight epilogue warps in two teams (a team = four TMEM lane groups, alternate tiles; a thread owns the 32 channels of one
//     pixel) add bias, apply leaky_relu, split to fp16 hi/lo and write two 256-bit stores per plane.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kBuilderWarps = 8, kEpiWarps = 8;
constexpr int kTeams = kEpiWarps / 4;                 // epilogue teams (one warp per TMEM lane group each) on alternate tiles
constexpr int kEpiWarp0 = kBuilderWarps;              // warps 0-7 builders (the first warp of a group also issues its MMAs), 8-15 epilogue
constexpr int kThreads = (kBuilderWarps + kEpiWarps) * 32;       // 512: four warps per scheduler, 128 registers per thread
constexpr int CO = 32;                                // output channels
constexpr int PH = 2 * kTileRows + 3, PW = 2 * kTileCols + 3;    // 35 x 19 input pixels
constexpr int PPITCH = 20;                            // fp16 per row of a converted plane (19 used + the zero slot 19; 10 words: the
                                                      // im2col loads of a warp -- 4 row pairs, 4 plane rows apart -- fall into
                                                      // 32 distinct banks)
constexpr int PPLANE = PH * PPITCH;                   // fp16 per (colour plane, half)
constexpr int RAW_PITCH = 64;                         // bytes per staged raw row: 19 pixels * 3 bytes + up to 3 bytes of misalignment
constexpr int A_TILE = kTileM * 64;                   // 8 KB: 128 rows x 32 fp16
constexpr int STAGE_BYTES = 2 * A_TILE;               // hi + lo
constexpr int GROUPS = 4;                             // builder groups (2 warps each) working on different items
constexpr int GTHREADS = kBuilderWarps * 32 / GROUPS;  // 64
constexpr int STAGES = 2 * GROUPS;                    // two A stages per group, alternating over the group's (item, plane) sequence
constexpr int PATCH_BYTES = 3 * 2 * PPLANE * 2;       // 3 colour planes x (hi, lo) x fp16 = 8 400 B
constexpr int RAW_BYTES = PH * RAW_PITCH;             // 2 240 B
constexpr int GROUP_BYTES = (PATCH_BYTES + RAW_BYTES + 15) / 16 * 16;
constexpr int W_TILE = CO * 64;                       // 2 KB: 32 rows x 32 fp16
constexpr int W_SET = 2 * W_TILE;                     // [W_hi | W_lo]
constexpr int SLOT_COLS = 2 * CO, SLOTS = 8, TMEM_COLS = 512;
constexpr int PATCH_OFF = STAGES * STAGE_BYTES;
constexpr int W_OFF = (PATCH_OFF + GROUPS * GROUP_BYTES + 1023) / 1024 * 1024;   // swizzle patterns are functions of the absolute address
constexpr int BAR_OFF = (W_OFF + 2 * W_SET + 1023) / 1024 * 1024;
constexpr int SMEM_BYTES = BAR_OFF + 512 + 2 * CO * 4 + 1024;
constexpr int RAW_WORDS = PH * (RAW_PITCH / 4);       // 560 words per item
constexpr int RAW_PER = (RAW_WORDS + GTHREADS - 1) / GTHREADS;   // 9 per thread
constexpr int PPAIRS_ROW = (PW + 1) / 2;                         // 10 column pairs per patch row
constexpr int PPER = (PH * PPAIRS_ROW + GTHREADS - 1) / GTHREADS;   // 6 pixel pairs per thread

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

// The sequence of work items of one CTA: item j of the CTA is global item blockIdx.x + j * gridDim.x = (image n, tile).
// Group j % GROUPS builds it; its three planes are the group's local (item, plane) numbers 3 * (j / GROUPS) + plane, which
// alternate between the group's two A stages / TMEM slots.
struct ItemIter {
  int n, ty, tx;       // current item
  int sn, sy, sx;      // stride decomposed the same way
  __device__ __forceinline__ void init(int t0, int stride, int tiles_x, int tiles_y) {
    const int tpi = tiles_x * tiles_y;
    n = t0 / tpi; int r = t0 - n * tpi; ty = r / tiles_x; tx = r - ty * tiles_x;
    sn = stride / tpi; r = stride - sn * tpi; sy = r / tiles_x; sx = r - sy * tiles_x;
  }
  __device__ __forceinline__ void next(int tiles_x, int tiles_y) {
    tx += sx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    ty += sy; if (ty >= tiles_y) { ty -= tiles_y; ++n; }
    n += sn;
  }
};

// A word that straddles an end of the caller's buffer, assembled from its in-range bytes (at most two words per batch).
__device__ __noinline__ uint32_t load_edge_word(const uint8_t* abase, int a, int lo, int hi) {
  uint32_t v = 0u;
#pragma unroll 1
  for (int bb = 0; bb < 4; ++bb)
    if (a + bb >= lo && a + bb < hi) v |= (uint32_t)__ldg(abase + a + bb) << (8 * bb);
  return v;
}

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GTHREADS) : "memory"); }

template <int IN_KIND /*0 rgb u8, 1 f32 planes*/>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv1(const __grid_constant__ TcConv1Params prm, int tiles_x, int tiles_y, int num_items, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  const WaitCtx wc{error_flag, prm.wait_timeout, 1};
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* w_base = smem + W_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty_bar = bars;                   // [STAGES]  MMAs done reading a stage -> builders
  uint64_t* slot_full = bars + STAGES;          // [SLOTS]   a tile's stage and TMEM slot share index and phase
  uint64_t* slot_empty = slot_full + SLOTS;     // [SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + SLOTS);
  float* bias_s = reinterpret_cast<float*>(smem + BAR_OFF + 512);      // [2][32]
  static_assert((STAGES + 2 * SLOTS) * 8 + 4 <= 512, "barrier area too small");
  static_assert(STAGES == SLOTS, "a tile's stage and TMEM slot share index and phase");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = prm.N, H = prm.H, W = prm.W, Ho = prm.Ho, Wo = prm.Wo;
  (void)num_items;
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * CO; i += kThreads) bias_s[i] = prm.bias[i] * ACT_SCALE;   // the epilogue works on scaled values
  // weights: [set][W_hi | W_lo], each [32 channels][32 K slots] fp16 K-major, SWIZZLE_64B; 16-byte chunks copied by all threads
  for (int i = threadIdx.x; i < 2 * 2 * CO * 4; i += kThreads) {
    const int j = i & 3, row = (i >> 2) & (CO - 1), part = (i >> 7) & 1, set = i >> 8;
    const uint4 v = *reinterpret_cast<const uint4*>((part ? prm.w_lo : prm.w_hi) + ((size_t)(set * CO + row) * 32 + j * 8));
    *reinterpret_cast<uint4*>(w_base + set * W_SET + part * W_TILE + sw64(row, j)) = v;
  }
  // the zero slot (column 19) and the unused columns of every converted plane stay zero for the whole kernel
  for (int i = threadIdx.x; i < GROUPS * GROUP_BYTES / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + PATCH_OFF)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                  // everything above only touched shared memory, TMEM and the (static) weights
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kBuilderWarps) {
    // ===================== builders: RGB patch -> three converted planes -> im2col tiles in the MMA layout, then the MMAs =====================
    const int g = warp / (kBuilderWarps / GROUPS);
    const int gt = threadIdx.x - g * GTHREADS;            // 0..63 inside the group
    uint8_t* gbase = smem + PATCH_OFF + g * GROUP_BYTES;
    __half* planes_s = reinterpret_cast<__half*>(gbase);                  // [3 colour planes][hi, lo][PH][PPITCH]
    uint8_t* raw_s = gbase + PATCH_BYTES;                                 // [PH][RAW_PITCH] staged RGB bytes
    // the patch pixel PAIRS this thread converts: i = gt + 64 q -> (row, column pair); fixed for the whole kernel.  A row has
    // ten pairs: columns (0,1) .. (18,19), column 19 being the zero slot of the K layout
    int16_t prc[PPER];                                     // row | column pair << 8
#pragma unroll
    for (int q = 0; q < PPER; ++q) {
      const int i = gt + q * GTHREADS;
      const int pr = i / PPAIRS_ROW, pcp = i - pr * PPAIRS_ROW;
      prc[q] = (int16_t)(pr | (pcp << 8));
    }
    ItemIter it;
    it.init(blockIdx.x + g * gridDim.x, GROUPS * gridDim.x, tiles_x, tiles_y);
    // Raw RGB words of an item: row r of the patch starts at byte (n, iy0 + r, ix0) of the image; the group stages the
    // aligned words that cover its 57 bytes.  Loads are unconditional from an always-valid (clamped) address, so all of
    // them are in flight together; rows / columns outside the image are zeroed at conversion (= TF SAME padding).  The
    // words of item k+1 are requested before item k is built, so their latency hides behind the im2col work.
    // Byte offsets are 32-bit and relative to `abase`, the aligned word at or below the start of the batch (the launcher
    // refuses batches of 2 GiB or more; the host side never forms them).
    const int mis0 = (int)(reinterpret_cast<uintptr_t>(prm.rgb) & 3);
    const uint8_t* abase = prm.rgb - mis0;
    const int total = N * H * W * 3;
    const int last_w = ((mis0 + total + 3) & ~3) - 4;                      // last aligned word that holds image bytes
    // the first / last word may straddle the ends of the caller's buffer (a batch that does not start or end on a 4-byte
    // boundary): those two words are assembled from in-range byte loads, every other word is one aligned 32-bit load
    const int full_lo = mis0 ? 4 : 0, full_hi = ((mis0 + total) & ~3) - 4;
    const int my_r = gt >> 4, my_wd4 = (gt & 15) * 4;                      // word i = gt + 64 q of the staged patch: row my_r + 4 q
    uint32_t rawreg[RAW_PER];
    auto load_raw = [&](const ItemIter& t) {
      const int iy0 = t.ty * kTileRows * 2 - prm.pad_t, ix0 = t.tx * kTileCols * 2 - prm.pad_l;
      const int nH = t.n * H;
#pragma unroll
      for (int q = 0; q < RAW_PER; ++q) {
        const int r = my_r + 4 * q;
        int iy = iy0 + r;
        iy = min(max(iy, 0), H - 1);                       // rows outside the image: any valid row (masked at conversion)
        const int row0 = mis0 + ((nH + iy) * W + ix0) * 3; // may be negative (left padding of the first row)
        const int a = min(max((row0 & ~3) + my_wd4, 0), last_w);
        uint32_t v = 0u;
        if (q < RAW_PER - 1 || r < PH) {
          if (a >= full_lo && a <= full_hi) {
            v = __ldg(reinterpret_cast<const uint32_t*>(abase + a));
          } else {
            v = load_edge_word(abase, a, mis0, mis0 + total);
          }
        }
        rawreg[q] = v;
      }
    };
    int j_local = 0;                                       // this group's item counter
    if (IN_KIND == 0 && it.n < N) load_raw(it);
    for (; it.n < N; ++j_local) {
      const int n = it.n;
      const int iy0 = it.ty * kTileRows * 2 - prm.pad_t, ix0 = it.tx * kTileCols * 2 - prm.pad_l;
      const bool interior = iy0 >= 0 && iy0 + PH <= H && ix0 >= 0 && ix0 + PW <= W;
      group_barrier(g);                          // the previous item of this group no longer reads the planes / raw bytes
      if (IN_KIND == 0) {
#pragma unroll
        for (int q = 0; q < RAW_PER; ++q) {
          const int i = gt + q * GTHREADS;
          if (q < RAW_PER - 1 || i < RAW_WORDS) reinterpret_cast<uint32_t*>(raw_s)[i] = rawreg[q];
        }
        group_barrier(g);
      }
      // ---- conversion: every pixel normalised once, projected onto Y, Cb, Cr, split once per plane; two neighbouring pixels
      //      per step, so that their fp16 halves are stored as one 32-bit word per plane ----
      // byte address of patch pixel (0, 0); it may lie outside the buffer (padding), only its low two bits are used
      const uint32_t img0_lo = (uint32_t)(mis0 + ((n * H + iy0) * W + ix0) * 3);
      const uint32_t row_bytes = (uint32_t)W * 3u;
      uint32_t* planes_w = reinterpret_cast<uint32_t*>(planes_s);
#pragma unroll
      for (int q = 0; q < PPER; ++q) {
        if (q < PPER - 1 || gt + q * GTHREADS < PH * PPAIRS_ROW) {
          const int pr = prc[q] & 0xff, pc = 2 * (prc[q] >> 8);
          const bool okr = interior || (unsigned)(iy0 + pr) < (unsigned)H;
          const bool ok0 = okr && (interior || (unsigned)(ix0 + pc) < (unsigned)W);
          const bool ok1 = okr && pc + 1 < PW && (interior || (unsigned)(ix0 + pc + 1) < (unsigned)W);
          float v0[3], v1[3];
          if (IN_KIND == 0) {
            int iy = iy0 + pr;
            iy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
            // the staged row starts at the aligned word below byte (iy, ix0); this pixel sits (row0 & 3) + 3 pc bytes in
            // (a clamped row or word holds other bytes, but only for pixels that are masked out below)
            const uint32_t mis = (img0_lo + (uint32_t)(iy - iy0) * row_bytes) & 3u;
            // the six bytes of the pixel pair: three aligned words, funnel-shifted by the byte misalignment
            const uint32_t boff = (uint32_t)(pr * RAW_PITCH + 3 * pc) + mis;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(raw_s + (boff & ~3u));
            const uint32_t sh = (boff & 3u) * 8u;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
            const uint32_t p03 = __funnelshift_r(w0, w1, sh), p45 = __funnelshift_r(w1, w2, sh);
            // x.astype(float32)/255, then (t0*k0 + t1*k1) + t2*k2 with separate roundings, then + offset
            const float r0 = div255(p03 & 0xffu), g0 = div255((p03 >> 8) & 0xffu), b0 = div255((p03 >> 16) & 0xffu);
            const float r1 = div255(p03 >> 24), g1 = div255(p45 & 0xffu), b1 = div255((p45 >> 8) & 0xffu);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
              v0[pl] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, prm.cc.k[pl][0]), __fmul_rn(g0, prm.cc.k[pl][1])),
                                           __fmul_rn(b0, prm.cc.k[pl][2])), prm.cc.off[pl]);
              v1[pl] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1, prm.cc.k[pl][0]), __fmul_rn(g1, prm.cc.k[pl][1])),
                                           __fmul_rn(b1, prm.cc.k[pl][2])), prm.cc.off[pl]);
            }
          } else {
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
              const float* pf = prm.planes + (((size_t)(pl * N + n) * H + (iy0 + pr)) * W + (ix0 + pc));
              v0[pl] = ok0 ? pf[0] : 0.0f;
              v1[pl] = ok1 ? pf[1] : 0.0f;
            }
          }
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            uint32_t hi2, lo2;                   // split once per input pixel; every tap that uses it copies the halves
            split2_f32(ok0 ? v0[pl] : 0.0f, ok1 ? v1[pl] : 0.0f, hi2, lo2);
            planes_w[((pl * 2 + 0) * PPLANE + pr * PPITCH + pc) >> 1] = hi2;
            planes_w[((pl * 2 + 1) * PPLANE + pr * PPITCH + pc) >> 1] = lo2;
          }
        }
      }
      it.next(tiles_x, tiles_y);
      if (IN_KIND == 0 && it.n < N) load_raw(it);          // next item of this group
      group_barrier(g);                          // the three planes are complete
      // ---- per plane: im2col rows into the group's next A stage, then its MMAs ----
#pragma unroll 1
      for (int pl = 0; pl < 3; ++pl) {
        const int lc = 3 * j_local + pl;                   // the group's (item, plane) counter
        const int stage = g * 2 + (lc & 1);
        const uint32_t phase = (uint32_t)(lc >> 1) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1, wc, 1);    // the MMAs that read this stage two planes ago are done
        uint8_t* a_hi = stage_base + stage * STAGE_BYTES;
        uint8_t* a_lo = a_hi + A_TILE;
        const uint32_t* ph_w = reinterpret_cast<const uint32_t*>(planes_s + (pl * 2 + 0) * PPLANE);
        const uint32_t* pl_w = reinterpret_cast<const uint32_t*>(planes_s + (pl * 2 + 1) * PPLANE);
        {
          // a thread builds the rows of two vertically neighbouring output pixels (r, c) and (r + 1, c), r even: their 5 x 6
          // input windows (input rows 2r .. 2r+4 and 2r+2 .. 2r+6, columns 2c .. 2c+5) share three of seven rows, so 21 aligned
          // 32-bit loads per half serve both.  K slot 6 kh + s = input pixel (2r + kh, 2c + s).  Eight consecutive threads
          // hold one tile row (c = 0..7): their 16-byte stores fall on the eight distinct chunks of a 128-byte swizzle atom
          // pair, and a warp's loads (4 row pairs x 8 columns, 40 words apart) on 32 distinct banks.
          const int rp = gt >> 3, c = gt & 7;
          const int m = rp * 2 * kTileCols + c;
          const int w0 = (4 * rp) * (PPITCH / 2) + c;                // word index of input pixel (4 rp, 2c)
          uint32_t hw[7][3], lw[7][3];
#pragma unroll
          for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              hw[k][s] = ph_w[w0 + k * (PPITCH / 2) + s];
              lw[k][s] = pl_w[w0 + k * (PPITCH / 2) + s];
            }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // words 4j .. 4j+3 of the row: word q = (kernel row q / 3, pixel pair q % 3); word 15 = K slots 30, 31 = zero
              uint32_t vh[4], vl[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int wq = 4 * j + q;
                vh[q] = wq < 15 ? hw[2 * e + wq / 3][wq % 3] : 0u;
                vl[q] = wq < 15 ? lw[2 * e + wq / 3][wq % 3] : 0u;
              }
              *reinterpret_cast<uint4*>(a_hi + sw64(m + e * kTileCols, j)) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
              *reinterpret_cast<uint4*>(a_lo + sw64(m + e * kTileCols, j)) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to tcgen05.mma
        group_barrier(g);                        // both warps of the group have written their rows
        if ((warp & (kBuilderWarps / GROUPS - 1)) == 0) {
          // ---- MMA issue: A_hi x [W_hi | W_lo] and A_lo x W_hi, two k-steps, into TMEM slot `stage`
          mbar_wait(&slot_empty[stage], phase ^ 1, wc, 2);
          tc_fence_after();
          if (elect_one()) {
            const int set = pl == 0 ? 0 : 1;
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
    }
  } else {
    // ===================== epilogue warps =====================
    // Teams of four warps (one warp per TMEM lane group) share the (item, plane) tiles of the CTA's sequence; a thread owns
    // ALL 32 channels of one pixel, so the per-tile bookkeeping (barrier wait, slot arithmetic, pixel address) is paid once
    // per 32 values, and while one team waits for its TMEM loads the others compute.  A TMEM slot always belongs to the same
    // team (slot % kTeams): every use of a slot is then drained in order by one team, so a parity wait can never be
    // satisfied by the slot's previous-but-one use.
    const int ew = warp - kEpiWarp0;
    const int lg = warp & 3;                  // TMEM lane group
    const int team = ew >> 2;
    const int m = lg * 32 + lane;
    const int my = m >> 3, mx = m & 7;
    ItemIter it;
    it.init(blockIdx.x, gridDim.x, tiles_x, tiles_y);
    for (int j = 0; it.n < N; ++j, it.next(tiles_x, tiles_y)) {      // item j of the CTA, built by group j % GROUPS
     for (int pl = 0; pl < 3; ++pl) {
      const int lc = 3 * (j / GROUPS) + pl;    // the building group's (item, plane) counter
      const int slot = (j % GROUPS) * 2 + (lc & 1);
      if (slot % kTeams != team) continue;
      mbar_wait(&slot_full[slot], (uint32_t)(lc >> 1) & 1u, wc, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * SLOT_COLS;
      uint32_t vm[CO], vc[CO];
      tmem_ld32_nowait(taddr, vm);
      tmem_ld32_nowait(taddr + CO, vc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[slot]);
      // values are kept multiplied by ACT_SCALE (a power of two: bias add, leaky_relu and the hi/lo split commute with it
      // bit for bit), which saves the scaling multiplies of the split
      const int set = pl == 0 ? 0 : 1;
      const float inv16 = prm.inv_scale[set] * ACT_SCALE;
      const float4* bs4 = reinterpret_cast<const float4*>(bias_s + set * CO);
      uint32_t h[CO / 2], l[CO / 2];
#pragma unroll
      for (int i = 0; i < CO; i += 4) {
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
      // 32 channels = 64 bytes per fp16 plane: two 256-bit stores each (two full sectors per thread)
      const int y = it.ty * kTileRows + my, x = it.tx * kTileCols + mx;
      if (y < Ho && x < Wo) {
        const size_t o = (((size_t)(pl * N + it.n) * prm.Hs + y) * prm.Ws + x) * CO;
        st_global_v8(prm.out_hi + o, h);
        st_global_v8(prm.out_hi + o + 16, h + 8);
        st_global_v8(prm.out_lo + o, l);
        st_global_v8(prm.out_lo + o + 16, l + 8);
      }
     }
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
  const long long items = (long long)tiles_x * tiles_y * prm.N;      // one item = one tile of one image, all three planes
  if (items <= 0 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (prm.rgb && (long long)prm.N * prm.H * prm.W * 3 > 0x7ffffff0LL) return cudaErrorInvalidValue;   // 32-bit byte offsets in the builders
  const int grid = items < num_sms ? (int)items : num_sms;
  return launch_kernel(prm.rgb ? k_tc_conv1<0> : k_tc_conv1<1>, dim3(grid), dim3(kThreads), SMEM_BYTES, stream, true, prm, tiles_x, tiles_y,
                       (int)items, error_flag);
}

}  // namespace nnic
