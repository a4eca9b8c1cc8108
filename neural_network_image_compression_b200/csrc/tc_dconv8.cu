// dconv8 + colour inverse + uint8 pack on the tensor cores.
//
// Reference: BaseDecoder.dconv8 + clip (decoder.py:17,31-32): Conv2DTranspose(1,5,2,'SAME',leaky_relu),
//   out[2i+a-1, 2j+b-1] += x[i,j,ci] * K[a,b,0,ci]          (64 input channels, ONE output channel)
// then Decoder.__call__ lines 45-48: convert_to_rgb (utils.py:70-72), clip(0,1), round(*255), uint8.
//
// With a single output channel the layer is not a GEMM over output pixels, but it is one over INPUT pixels:
//   R[q, t] = sum_ci x[q, ci] * K[t, ci]      q = input pixel, t = a*5+b (25 taps, padded to N = 32)
// followed by a gather out[y, x] = sum over the (at most 9) taps that reach it of R[neighbour, t].
// One work item = 14 x 6 input pixels of one image (+ one pixel of halo = one 16 x 8 = 128-pixel A tile)
// -> 28 x 12 output pixels, all three colour planes (plane 0 with the 'Y' weights, planes 1/2 with 'CbCr').
// Per plane: one TMA load of the tile (hi, lo), 4 k-steps x 2 MMAs (A_hi x [W_hi|W_lo] with N = 64, A_lo x W_hi
// with N = 32), accumulator [main 32 | corr 32] in one of eight TMEM slots.  Four epilogue warps read R, exchange
// it through shared memory, gather the outputs in a fixed order, and after the third plane apply the colour
// transform and store packed RGB bytes.  The epilogue is the long pole (a few hundred dependent instructions per
// thread and item on one warp per scheduler), so two such teams of four warps work on alternate items.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kTeams = 2;                     // epilogue teams (4 warps = 128 TMEM lanes each) on alternate items
constexpr int kThreads = (2 + 4 * kTeams) * 32;   // warp 0 TMA, warp 1 MMA issuer, then the epilogue warps
constexpr int NT = 32;                        // taps padded to the MMA N granularity
constexpr int IH = kTileRows - 2, IW = kTileCols - 2;   // interior input pixels per item: 14 x 6
constexpr int OH = 2 * IH, OW = 2 * IW;       // output pixels per item: 28 x 12
constexpr int NOUT = OH * OW;                 // 336
constexpr int A_BYTES = kTileM * 128;         // 16 KB per hi / lo tile
constexpr int STAGE_BYTES = 2 * A_BYTES;
constexpr int STAGES = 4;
constexpr int W_TILE = NT * 128;              // 4 KB: [32 taps][64 ci] fp16
constexpr int W_SET = 2 * W_TILE;             // [W_hi | W_lo]
constexpr int SLOT_COLS = 2 * NT, SLOTS = 8, TMEM_COLS = 512;
constexpr int W_OFF = STAGES * STAGE_BYTES;
constexpr int RESP_OFF = W_OFF + 2 * W_SET;                   // float [team][2 buffers][25][128]
constexpr int RESP_TEAM = 2 * 25 * 128 * 4;
constexpr int RGB_OFF = RESP_OFF + kTeams * RESP_TEAM;        // uint8 [team][28][36]
constexpr int RGB_TEAM = (OH * OW * 3 + 15) / 16 * 16;
constexpr int BAR_OFF = RGB_OFF + kTeams * RGB_TEAM;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;

__device__ __forceinline__ void epi_barrier(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
k_tc_dconv8(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
            const __grid_constant__ TcDconv8Params prm, int tiles_x, int tiles_y, int num_items, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  const WaitCtx wc{error_flag, prm.wait_timeout, 12};
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* w_base = smem + W_OFF;                               // [set][W_hi | W_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* slot_full = bars + 2 * STAGES;      // [SLOTS]
  uint64_t* slot_empty = slot_full + SLOTS;     // [SLOTS]
  uint64_t* w_bar = slot_empty + SLOTS;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  static_assert((2 * STAGES + 2 * SLOTS + 1) * 8 + 4 <= 256, "barrier area too small");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = prm.N, Hi = prm.Hi, Wi = prm.Wi, Ho = 2 * prm.Hi, Wo = 2 * prm.Wi;
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], 4); }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_image = tiles_x * tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {                       // both weight sets stay resident: 16 KB
      mbar_expect_tx(w_bar, 2 * W_SET);
      for (int set = 0; set < 2; ++set) {
        tma_load_2d(&map_w_hi, w_base + set * W_SET, w_bar, 0, set * NT);
        tma_load_2d(&map_w_lo, w_base + set * W_SET + W_TILE, w_bar, 0, set * NT);
      }
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int txy = it % tiles_per_image, n = it / tiles_per_image;
      const int Y0 = (txy / tiles_x) * IH - 1, X0 = (txy % tiles_x) * IW - 1;
      for (int plane = 0; plane < 3; ++plane) {
        mbar_wait(&empty_bar[stage], phase ^ 1, wc, 1);
        if (elect_one()) {
          uint8_t* sb = stage_base + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], prm.fast ? A_BYTES : STAGE_BYTES);
          tma_load_5d(&map_a_hi, sb, &full_bar[stage], 0, X0, 0, Y0, plane * N + n);
          if (!prm.fast) tma_load_5d(&map_a_lo, sb + A_BYTES, &full_bar[stage], 0, X0, 0, Y0, plane * N + n);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_wide = make_idesc(2 * NT);
    constexpr uint32_t idesc_narrow = make_idesc(NT);
    const uint32_t stage_u32 = smem_u32(stage_base), w_u32 = smem_u32(w_base);
    mbar_wait(w_bar, 0, wc, 2);
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      for (int plane = 0; plane < 3; ++plane) {
        mbar_wait(&slot_empty[slot], slot_phase ^ 1, wc, 3);
        mbar_wait(&full_bar[stage], phase, wc, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * SLOT_COLS;
        const uint64_t a_hi = make_smem_desc<128>(stage_u32 + stage * STAGE_BYTES);
        const uint64_t a_lo = a_hi + (uint64_t)(A_BYTES >> 4);
        const uint64_t w_hl = make_smem_desc<128>(w_u32 + (plane == 0 ? 0 : W_SET));   // [W_hi | W_lo]: 64 rows
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, ks ? 1u : 0u);     // A_hi x [W_hi | W_lo]
            if (!prm.fast) umma_f16(d_tmem + NT, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&slot_full[slot]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // thread = haloed input pixel (r, c) of the 16 x 8 tile.  It publishes its 25 tap responses, and if it is an
    // interior pixel it gathers the 2 x 2 output block at (2(r-1), 2(c-1)): 9 + 6 + 6 + 4 = 25 neighbour
    // responses at compile-time offsets.  The three planes' outputs stay in registers until the colour step.
    const int lg = warp & 3;
    const int q = lg * 32 + lane;
    const int team = (warp - 2) >> 2;
    const int et = ((warp - 2) & 3) * 32 + lane;    // 0..127 among the team's threads
    float (*resp_s)[128] = reinterpret_cast<float (*)[128]>(smem + RESP_OFF + team * RESP_TEAM);
    uint8_t* rgb_s = smem + RGB_OFF + team * RGB_TEAM;
    const int r = q >> 3, c = q & 7;
    const bool interior = r >= 1 && r <= IH && c >= 1 && c <= IW;
    int rbuf = 0;
    int k_item = 0;                            // this CTA's item counter: item k uses TMEM slots 3k .. 3k+2 (mod SLOTS)
    for (int it = blockIdx.x; it < num_items; it += gridDim.x, ++k_item) {
      if ((k_item & (kTeams - 1)) != team) continue;
      const int txy = it % tiles_per_image, n = it / tiles_per_image;
      const int ty = txy / tiles_x, tx = txy % tiles_x;
      float outv[3][4];                        // [plane][py*2+px]
#pragma unroll
      for (int plane = 0; plane < 3; ++plane) {
        const int set = plane == 0 ? 0 : 1;
        const int cnt = 3 * k_item + plane;
        const int slot = cnt & (SLOTS - 1);
        const uint32_t slot_phase = (cnt / SLOTS) & 1;
        mbar_wait(&slot_full[slot], slot_phase, wc, 5);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * SLOT_COLS;
        uint32_t vm[NT], vc[NT];
        tmem_ld32_nowait(taddr, vm);
        tmem_ld32_nowait(taddr + NT, vc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[slot]);
        const float inv_scale = prm.inv_scale[set];       // a power of two: scaling the gathered sum is bit-identical to scaling its terms
        float (*resp)[128] = resp_s + rbuf * 25;
#pragma unroll
        for (int t = 0; t < 25; ++t)
          resp[t][q] = __fadd_rn(__uint_as_float(vm[t]), __uint_as_float(vc[t]));
        epi_barrier(team);                     // also orders the previous use of the other buffer (see below)
        if (interior) {
          const float bias = prm.bias[set];
#pragma unroll
          for (int py = 0; py < 2; ++py) {
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              float acc = 0.0f;
              // taps a = (py+1) mod 2 (+2, +4): input row = r + (py+1-a)/2
#pragma unroll
              for (int ta = (py + 1) & 1; ta < 5; ta += 2) {
#pragma unroll
                for (int tb = (px + 1) & 1; tb < 5; tb += 2)
                  acc = __fadd_rn(acc, resp[ta * 5 + tb][q + ((py + 1 - ta) / 2) * kTileCols + (px + 1 - tb) / 2]);
              }
              const float v = leaky(__fadd_rn(__fmul_rn(acc, inv_scale), bias));
              outv[plane][py * 2 + px] = fminf(fmaxf(v, 0.0f), 1.0f);      // decoder.py:32
            }
          }
        }
        // no second barrier: the next plane writes the OTHER buffer, and nobody can pass the next barrier
        // before every thread has finished this gather
        rbuf ^= 1;
      }
      const int oy0 = ty * OH, ox0 = tx * OW;
      if (interior) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int ry = 2 * (r - 1) + (k4 >> 1), rx = 2 * (c - 1) + (k4 & 1);
          const int oy = oy0 + ry, ox = ox0 + rx;
          const float y = outv[0][k4], cb = outv[1][k4], cr = outv[2][k4];
          const bool ok = oy < Ho && ox < Wo;
          if (prm.planes_out && ok) {
            const size_t plane_sz = (size_t)N * Ho * Wo;
            const size_t g = ((size_t)n * Ho + oy) * Wo + ox;
            prm.planes_out[g] = y;
            prm.planes_out[plane_sz + g] = cb;
            prm.planes_out[2 * plane_sz + g] = cr;
          }
          // convert_to_rgb: subtract the offsets, project with the inverse kernel, clip (decoder.py:45-46)
          const float t0 = __fsub_rn(y, prm.cc.off[0]), t1 = __fsub_rn(cb, prm.cc.off[1]), t2 = __fsub_rn(cr, prm.cc.off[2]);
          float ch[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(t0, prm.cc.kinv[k][0]), __fmul_rn(t1, prm.cc.kinv[k][1])),
                                      __fmul_rn(t2, prm.cc.kinv[k][2]));
            ch[k] = fminf(fmaxf(v, 0.0f), 1.0f);
          }
          if (prm.prequant && ok) {
            const size_t g = (((size_t)n * Ho + oy) * Wo + ox) * 3;
            prm.prequant[g] = ch[0]; prm.prequant[g + 1] = ch[1]; prm.prequant[g + 2] = ch[2];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) rgb_s[ry * (OW * 3) + rx * 3 + k] = (uint8_t)rintf(__fmul_rn(ch[k], 255.0f));   // decoder.py:48
        }
      }
      epi_barrier(team);
      if (prm.rgb) {
        constexpr int ROWB = OW * 3;                            // 36 bytes per tile row
        const bool vec4 = (ox0 + OW <= Wo) && (Wo % 4 == 0);
        if (vec4) {
          for (int i = et; i < OH * (ROWB / 4); i += 128) {
            const int r_ = i / (ROWB / 4), w_ = i - r_ * (ROWB / 4);
            const int oy = oy0 + r_;
            if (oy < Ho)
              *reinterpret_cast<uint32_t*>(prm.rgb + (((size_t)n * Ho + oy) * Wo + ox0) * 3 + w_ * 4) =
                  *reinterpret_cast<const uint32_t*>(&rgb_s[r_ * ROWB + w_ * 4]);
          }
        } else {
          for (int i = et; i < OH * ROWB; i += 128) {
            const int r_ = i / ROWB, b = i - r_ * ROWB;
            const int oy = oy0 + r_, ox = ox0 + b / 3;
            if (oy < Ho && ox < Wo) prm.rgb[(((size_t)n * Ho + oy) * Wo) * 3 + (size_t)ox0 * 3 + b] = rgb_s[r_ * ROWB + b];
          }
        }
      }
      epi_barrier(team);                     // rgb_s is free for the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- dconv8 behind the fused dconv7 (tc_conv_patch.cu, FUSE8): gather of the tap responses + colour inverse + pack ----------
// R [3N][tile][25 taps][4 phases][128] fp32 holds, for every dconv7 OUTPUT pixel (y, x) = (2 iy + py, 2 ix + px) and tap t = 5a + b,
// the response sum_ci x[y, x, ci] K8[a, b, 0, ci] (unscaled accumulator); (iy, ix) runs over dconv7's Hp x Wp INPUT grid, stored in
// its 16 x 8-pixel tiles (row-major tile order, pixel m = 8 (iy % 16) + ix % 8 inside a tile), the order dconv7's epilogue
// produces them in.  Conv2DTranspose(1, 5, 2, 'SAME') (decoder.py:17):
//   out[2y + a - 1, 2x + b - 1] += R[y, x, 5a + b]
// so every response feeds exactly one output pixel and the layer is a permutation-sum over R.  One block = one tile of one image,
// one thread = one final row (4 pixels; RY = 0..3 = blockIdx.y, so the tap pattern is uniform in a block) of the 4 x 4 block of
// final pixels under input pixel (iy, ix): 20 (even rows) or 30 (odd rows) 4-byte loads per colour plane, neighbouring threads on
// neighbouring addresses, in the summation order of k_tc_dconv8's gather (taps a ascending, then b ascending), then bias, leaky,
// clip (decoder.py:31-32), convert_to_rgb, clip, round(*255) (decoder.py:45-48) and one 12-byte run of RGB.
template <int RY>
__device__ __forceinline__ void dconv8_gather_row(const float* __restrict__ R, const TcDconv8Params& prm, int Hp, int Wp, int tiles_x,
                                                  int tiles_per_plane, int n, int iy, int ix, bool rgb_aligned) {
  constexpr int py = RY >> 1, fy = RY & 1;
  constexpr int TILE_FLOATS = 25 * 4 * kTileM;
  const int N = prm.N;
  const int Ho = 4 * Hp, Wo = 4 * Wp;
  // float offset of pixel (iy + dy, ix + dx) inside a plane's R block, or -1 outside the plane (contributes zero = SAME padding)
  int noff[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int y = iy + dy, x = ix + dx;
      const bool ok = y >= 0 && y < Hp && x >= 0 && x < Wp;
      noff[dy + 1][dx + 1] = ok ? ((y >> 4) * tiles_x + (x >> 3)) * TILE_FLOATS + (y & 15) * kTileCols + (x & 7) : -1;
    }
  float outv[3][4];                                 // [colour plane][final column 0..3]
#pragma unroll
  for (int plane = 0; plane < 3; ++plane) {
    const int set = plane == 0 ? 0 : 1;
    const float inv_scale = prm.inv_scale[set], bias = prm.bias[set];
    const float* Rp = R + (size_t)(plane * N + n) * tiles_per_plane * TILE_FLOATS;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
#pragma unroll
      for (int fx = 0; fx < 2; ++fx) {
        float acc = 0.0f;
#pragma unroll
        for (int ta = (fy + 1) & 1; ta < 5; ta += 2) {
          const int yy = py + (fy + 1 - ta) / 2;            // dconv7 output row relative to 2 iy: -1 .. 2
          const int dy = yy < 0 ? -1 : (yy > 1 ? 1 : 0), npy = yy & 1;
#pragma unroll
          for (int tb = (fx + 1) & 1; tb < 5; tb += 2) {
            const int xx = px + (fx + 1 - tb) / 2;
            const int dx = xx < 0 ? -1 : (xx > 1 ? 1 : 0), npx = xx & 1;
            const int o = noff[dy + 1][dx + 1];
            const float r = o >= 0 ? __ldg(Rp + o + ((ta * 5 + tb) * 4 + npy * 2 + npx) * kTileM) : 0.0f;
            acc = __fadd_rn(acc, r);
          }
        }
        const float v = leaky(__fadd_rn(__fmul_rn(acc, inv_scale), bias));
        outv[plane][2 * px + fx] = fminf(fmaxf(v, 0.0f), 1.0f);      // decoder.py:32
      }
    }
  }
  const int oy = 4 * iy + RY, ox0 = 4 * ix;
  uint32_t packed[3] = {0u, 0u, 0u};
#pragma unroll
  for (int rx = 0; rx < 4; ++rx) {
    const float y = outv[0][rx], cb = outv[1][rx], cr = outv[2][rx];
    const size_t g = ((size_t)n * Ho + oy) * Wo + ox0 + rx;
    if (prm.planes_out) {
      const size_t plane_sz = (size_t)N * Ho * Wo;
      prm.planes_out[g] = y; prm.planes_out[plane_sz + g] = cb; prm.planes_out[2 * plane_sz + g] = cr;
    }
    // convert_to_rgb: subtract the offsets, project with the inverse kernel, clip (decoder.py:45-46)
    const float t0 = __fsub_rn(y, prm.cc.off[0]), t1 = __fsub_rn(cb, prm.cc.off[1]), t2 = __fsub_rn(cr, prm.cc.off[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float v = __fadd_rn(__fadd_rn(__fmul_rn(t0, prm.cc.kinv[k][0]), __fmul_rn(t1, prm.cc.kinv[k][1])),
                                __fmul_rn(t2, prm.cc.kinv[k][2]));
      const float c = fminf(fmaxf(v, 0.0f), 1.0f);
      if (prm.prequant) prm.prequant[g * 3 + k] = c;
      const uint32_t byte = (uint32_t)(uint8_t)rintf(__fmul_rn(c, 255.0f));                  // decoder.py:48
      const int b = rx * 3 + k;
      packed[b >> 2] |= byte << (8 * (b & 3));
    }
  }
  if (prm.rgb) {
    uint8_t* dst8 = prm.rgb + (((size_t)n * Ho + oy) * Wo + ox0) * 3;       // 12-byte runs: 4-byte aligned when the buffer is
    if (rgb_aligned) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(dst8);
      dst[0] = packed[0]; dst[1] = packed[1]; dst[2] = packed[2];
    } else {
#pragma unroll
      for (int b = 0; b < 12; ++b) dst8[b] = (uint8_t)(packed[b >> 2] >> (8 * (b & 3)));
    }
  }
}

__global__ void __launch_bounds__(kTileM)
k_dconv8_gather(const float* __restrict__ R, const __grid_constant__ TcDconv8Params prm, int Hp, int Wp, int tiles_x, int tiles_per_plane,
                bool rgb_aligned) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x / tiles_per_plane, txy = blockIdx.x - n * tiles_per_plane;
  const int iy = (txy / tiles_x) * kTileRows + (threadIdx.x >> 3), ix = (txy % tiles_x) * kTileCols + (threadIdx.x & 7);
  if (iy >= Hp || ix >= Wp) return;
  switch (blockIdx.y) {
    case 0: dconv8_gather_row<0>(R, prm, Hp, Wp, tiles_x, tiles_per_plane, n, iy, ix, rgb_aligned); break;
    case 1: dconv8_gather_row<1>(R, prm, Hp, Wp, tiles_x, tiles_per_plane, n, iy, ix, rgb_aligned); break;
    case 2: dconv8_gather_row<2>(R, prm, Hp, Wp, tiles_x, tiles_per_plane, n, iy, ix, rgb_aligned); break;
    default: dconv8_gather_row<3>(R, prm, Hp, Wp, tiles_x, tiles_per_plane, n, iy, ix, rgb_aligned); break;
  }
}

}  // namespace

cudaError_t launch_dconv8_gather(const float* R, const TcDconv8Params& prm, int Hp, int Wp, cudaStream_t stream) {
  const int tiles_x = (Wp + kTileCols - 1) / kTileCols, tiles_y = (Hp + kTileRows - 1) / kTileRows;
  const long long blocks = (long long)tiles_x * tiles_y * prm.N;
  // a plane's R block is indexed with 32-bit float offsets
  if (blocks <= 0 || blocks > 0x7fffffffLL || (long long)tiles_x * tiles_y * 12800 > 0x7fffffffLL) return cudaErrorInvalidValue;
  const bool rgb_aligned = (reinterpret_cast<uintptr_t>(prm.rgb) & 3) == 0;
  return launch_kernel(k_dconv8_gather, dim3((unsigned)blocks, 4), dim3(kTileM), 0, stream, true, R, prm, Hp, Wp, tiles_x, tiles_x * tiles_y,
                       rgb_aligned);
}

cudaError_t launch_tc_dconv8(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                             const CUtensorMap& w_lo, const TcDconv8Params& prm, int num_sms, int* error_flag,
                             cudaStream_t stream) {
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_dconv8, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const int tiles_x = (prm.Wi + IW - 1) / IW, tiles_y = (prm.Hi + IH - 1) / IH;
  const long long items = (long long)tiles_x * tiles_y * prm.N;
  if (items <= 0 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int grid = items < num_sms ? (int)items : num_sms;
  return launch_kernel(k_tc_dconv8, dim3(grid), dim3(kThreads), SMEM_BYTES, stream, true, a_hi, a_lo, w_hi, w_lo, prm, tiles_x, tiles_y, (int)items,
                       error_flag);
}

}  // namespace nnic
