// tcgen05 convolution, halo-patch variant, for the layers whose taps all lie in the 3x3 neighbourhood of the
// output pixel: conv3, conv4 (encoder.py:12-13), dconv5, dconv6 and the four output phases of dconv7
// (decoder.py:14-16).  64 input channels, 64 output channels.
//
// tc_conv.cu streams one [128 pixel x 64 channel] activation tile per tap through shared memory; measured
// with ncu, that kernel is bound by the shared-memory port (MMA operand reads ~125 B/clk plus TMA fills
// ~107 B/clk against a 128 B/clk port).  Here the activation patch of a tile (18 rows x 10 columns of
// pixels, 128 bytes each, SWIZZLE_128B) is loaded ONCE per work item and every tap reads it through a UMMA
// descriptor whose start address is shifted by whole pixel rows and whose 8-row-group stride is the patch
// pitch (10 pixels = 1280 bytes); tools/probe_tc.cu (probe A) shows the tensor core applies the 128-byte
// swizzle on absolute shared-memory addresses, so such descriptors read exactly the shifted rows.  Only the
// per-tap weight tiles (16 KB hi+lo) still stream, through an 8-deep ring.  For dconv7 one work item covers all
// four output phases of a tile: they share the same patch (25 taps in total).
//
// Arithmetic, accumulation chains, TMEM slot ring and epilogue are those of tc_conv.cu.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kThreads = 320;                 // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int COUT = 64;
constexpr int PH = kTileRows + 2, PW = kTileCols + 2;        // 18 x 10 pixels
constexpr int PATCH_TX = PH * PW * 128;                      // bytes one patch load brings: 23040
constexpr int PATCH_SLOT = (PATCH_TX + 1023) / 1024 * 1024;  // 23552
constexpr int SET_BYTES = 2 * PATCH_SLOT;                    // hi + lo
constexpr int NSETS = 2;
constexpr int W_TILE = COUT * 128;                           // 8192: one of hi / lo
constexpr int W_SLOT = 2 * W_TILE;
constexpr int WSLOTS = 8;
constexpr int SLOT_COLS = 2 * COUT, SLOTS = 4, TMEM_COLS = 512;
constexpr int BAR_OFF = NSETS * SET_BYTES + WSLOTS * W_SLOT;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 2 * COUT * 4 + 1024;
constexpr uint32_t A_SBO = PW * 128;                         // 8-row group stride of a tap view: one patch row

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv_patch(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ TcPatchParams prm, int tiles_x, int tiles_y, int num_items, int* error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* patch_base = smem;                                  // [NSETS][hi | lo]
  uint8_t* w_base = smem + NSETS * SET_BYTES;                  // [WSLOTS][W_hi | W_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* patch_full = bars;                 // [NSETS]
  uint64_t* patch_empty = bars + NSETS;        // [NSETS]
  uint64_t* w_full = bars + 2 * NSETS;         // [WSLOTS]
  uint64_t* w_empty = w_full + WSLOTS;         // [WSLOTS]
  uint64_t* slot_full = w_empty + WSLOTS;      // [SLOTS]
  uint64_t* slot_empty = slot_full + SLOTS;    // [SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + SLOTS);
  float* bias_s = reinterpret_cast<float*>(smem + BAR_OFF + 256);
  static_assert((2 * NSETS + 2 * WSLOTS + 2 * SLOTS) * 8 + 4 <= 256, "barrier area too small");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSETS; ++s) { mbar_init(&patch_full[s], 1); mbar_init(&patch_empty[s], 1); }
    for (int s = 0; s < WSLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
  }
  for (int i = threadIdx.x; i < 2 * COUT; i += kThreads) bias_s[i] = prm.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_plane = tiles_x * tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int pb = 0; uint32_t pphase = 0;
    int ws = 0; uint32_t wphase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int txy = it % tiles_per_plane;
      const int p = it / tiles_per_plane;
      const int Y0 = (txy / tiles_x) * kTileRows, X0 = (txy % tiles_x) * kTileCols;
      const int set = p < prm.n_split ? 0 : 1;
      mbar_wait(&patch_empty[pb], pphase ^ 1, error_flag, 1);
      if (elect_one()) {
        uint8_t* pbuf = patch_base + pb * SET_BYTES;
        mbar_expect_tx(&patch_full[pb], 2 * PATCH_TX);
        tma_load_5d(&map_a_hi, pbuf, &patch_full[pb], 0, X0 - 1, 0, Y0 - 1, p);
        tma_load_5d(&map_a_lo, pbuf + PATCH_SLOT, &patch_full[pb], 0, X0 - 1, 0, Y0 - 1, p);
      }
      __syncwarp();
      if (++pb == NSETS) { pb = 0; pphase ^= 1; }
      for (int j = 0; j < prm.njobs; ++j) {
        const int nsteps = prm.jobs[j].nsteps;
        for (int s = 0; s < nsteps; ++s) {
          const int wrow = set * prm.rows_per_set + prm.jobs[j].steps[s].w_row;
          mbar_wait(&w_empty[ws], wphase ^ 1, error_flag, 2);
          if (elect_one()) {
            uint8_t* wb = w_base + ws * W_SLOT;
            mbar_expect_tx(&w_full[ws], W_SLOT);
            tma_load_2d(&map_w_hi, wb, &w_full[ws], 0, wrow);
            tma_load_2d(&map_w_lo, wb + W_TILE, &w_full[ws], 0, wrow);
          }
          __syncwarp();
          if (++ws == WSLOTS) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_wide = make_idesc(2 * COUT);
    constexpr uint32_t idesc_narrow = make_idesc(COUT);
    const uint32_t patch_u32 = smem_u32(patch_base), w_u32 = smem_u32(w_base);
    int pb = 0; uint32_t pphase = 0;
    int ws = 0; uint32_t wphase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      mbar_wait(&patch_full[pb], pphase, error_flag, 3);
      const uint32_t pset = patch_u32 + pb * SET_BYTES;
      for (int j = 0; j < prm.njobs; ++j) {
        const int nsteps = prm.jobs[j].nsteps;
        const uint32_t chain_end_mask = prm.jobs[j].chain_end_mask;
        bool chain_start = true;
        uint32_t d_tmem = 0;
        for (int s = 0; s < nsteps; ++s) {
          const uint32_t a_off = prm.jobs[j].steps[s].a_off;
          const int chain_end = (chain_end_mask >> s) & 1u;
          if (chain_start) {
            mbar_wait(&slot_empty[slot], slot_phase ^ 1, error_flag, 4);
            d_tmem = tmem_base + slot * SLOT_COLS;
          }
          mbar_wait(&w_full[ws], wphase, error_flag, 5);
          tc_fence_after();
          const uint64_t a_hi = make_desc_sbo(pset + a_off, A_SBO);
          const uint64_t a_lo = a_hi + (uint64_t)(PATCH_SLOT >> 4);
          const uint64_t w_hl = make_desc_sbo(w_u32 + ws * W_SLOT, 1024);   // W_hi tile followed by the W_lo tile
          if (elect_one()) {
            umma_f16(d_tmem, a_hi, w_hl, idesc_wide, chain_start ? 0u : 1u);
            umma_f16(d_tmem + COUT, a_lo, w_hl, idesc_narrow, 1u);
#pragma unroll
            for (int ks = 1; ks < 4; ++ks) {
              umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, 1u);
              umma_f16(d_tmem + COUT, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
            }
            umma_commit(&w_empty[ws]);
            if (chain_end) umma_commit(&slot_full[slot]);
          }
          __syncwarp();
          if (++ws == WSLOTS) { ws = 0; wphase ^= 1; }
          chain_start = false;
          if (chain_end) {
            if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
            chain_start = true;
          }
        }
      }
      if (elect_one()) umma_commit(&patch_empty[pb]);      // every MMA of this work item has read the patch
      __syncwarp();
      if (++pb == NSETS) { pb = 0; pphase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    constexpr int HALF = COUT / 2;
    const int lg = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int m = lg * 32 + lane;
    const int r = m >> 3, c = m & 7;
    const int ch0 = hf * HALF;
    int slot = 0; uint32_t slot_phase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int txy = it % tiles_per_plane;
      const int p = it / tiles_per_plane;
      const int Y = (txy / tiles_x) * kTileRows + r, X = (txy % tiles_x) * kTileCols + c;
      const int set = p < prm.n_split ? 0 : 1;
      const float inv_scale = prm.inv_scale[set];
      const float* bs = bias_s + set * COUT + ch0;
      for (int j = 0; j < prm.njobs; ++j) {
        const int nchains = prm.jobs[j].nchains;
        const int oy = Y * prm.out_stride + prm.jobs[j].out_oy, ox = X * prm.out_stride + prm.jobs[j].out_ox;
        const bool valid = Y < prm.Hp && X < prm.Wp && oy < prm.Ho && ox < prm.Wo;
        float acc[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.0f;
        for (int ch = 0; ch < nchains; ++ch) {
          mbar_wait(&slot_full[slot], slot_phase, error_flag, 6);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * SLOT_COLS + ch0;
          uint32_t vm[HALF], vc[HALF];
          tmem_ld32_nowait(taddr, vm);
          tmem_ld32_nowait(taddr + COUT, vc);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_empty[slot]);
#pragma unroll
          for (int i = 0; i < HALF; ++i) acc[i] = __fadd_rn(acc[i], __fadd_rn(__uint_as_float(vm[i]), __uint_as_float(vc[i])));
          if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
        }
        if (valid) {
          const size_t pix = ((size_t)p * prm.Ho + oy) * prm.Wo + ox;
#pragma unroll
          for (int c0 = 0; c0 < HALF; c0 += 16) {
            float* v = acc + c0;
            const size_t o = pix * COUT + ch0 + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = leaky(__fadd_rn(v[i] * inv_scale, bs[c0 + i]));
            if (prm.res_hi) {
              __align__(16) __half rh[16], rl[16];
              *reinterpret_cast<uint4*>(rh) = *reinterpret_cast<const uint4*>(prm.res_hi + o);
              *reinterpret_cast<uint4*>(rh + 8) = *reinterpret_cast<const uint4*>(prm.res_hi + o + 8);
              *reinterpret_cast<uint4*>(rl) = *reinterpret_cast<const uint4*>(prm.res_lo + o);
              *reinterpret_cast<uint4*>(rl + 8) = *reinterpret_cast<const uint4*>(prm.res_lo + o + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(v[i], join_f32(rh[i], rl[i]));
            }
            if (prm.out_mode == TC_OUT_SPLIT) {
              __align__(16) __half h[16], l[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) split_f32(v[i], h[i], l[i]);
              *reinterpret_cast<uint4*>(prm.out_hi + o) = *reinterpret_cast<uint4*>(h);
              *reinterpret_cast<uint4*>(prm.out_hi + o + 8) = *reinterpret_cast<uint4*>(h + 8);
              *reinterpret_cast<uint4*>(prm.out_lo + o) = *reinterpret_cast<uint4*>(l);
              *reinterpret_cast<uint4*>(prm.out_lo + o + 8) = *reinterpret_cast<uint4*>(l + 8);
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(prm.out_f32 + o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace

uint32_t tc_patch_a_offset(int dy, int dx) { return (uint32_t)(((dy + 1) * PW + (dx + 1)) * 128); }

cudaError_t launch_tc_conv_patch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                 const CUtensorMap& w_lo, const TcPatchParams& prm, int num_sms, int* error_flag,
                                 cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_conv_patch, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles_x = (prm.Wp + kTileCols - 1) / kTileCols, tiles_y = (prm.Hp + kTileRows - 1) / kTileRows;
  const long long items = (long long)tiles_x * tiles_y * prm.P;
  if (items <= 0 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int grid = items < num_sms ? (int)items : num_sms;
  k_tc_conv_patch<<<grid, kThreads, SMEM_BYTES, stream>>>(a_hi, a_lo, w_hi, w_lo, prm, tiles_x, tiles_y, (int)items, error_flag);
  return cudaGetLastError();
}

}  // namespace nnic
