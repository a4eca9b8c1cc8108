// tcgen05 implicit-GEMM convolution for the eight GEMM-shaped layers of the codec
// (conv2/3/4/8, dconv1/5/6/7 -- reference tf2_0/src/encoder.py:11-17, decoder.py:10-16).
//
// GEMM view per output tile:  D[128 pixels, COUT] = sum over steps  A_step[128 pixels, K] x W_step[K, COUT]
//   * a tile is 16 rows x 8 columns of one output parity phase of one colour plane
//   * A_step is the input activation patch of the tap (a TMA box [K, 8, 1, 16, 1] out of a 5-D view of
//     the NHWC tensor; out-of-image pixels are zero-filled by the TMA unit = SAME padding)
//   * operands are fp16 hi/lo pairs: D = Ahi*Whi + (Ahi*Wlo + Alo*Whi).  The first two products are ONE
//     MMA against the stacked B operand [Whi | Wlo] (N = 2*COUT; measured: a 128xNx16 MMA from shared memory
//     costs max(48, N/2) cycles, so N = 128 runs at the pipe peak while three N = 64 MMAs would not)
//   * the tensor core truncates its fp32 accumulator at every MMA (measured, tools/probe_tc.cu), so a tile is
//     accumulated as several short chains, each in its own TMEM slot [main | correction]; the epilogue adds
//     the chains in registers with round-to-nearest fp32 adds
//   * epilogue: 2^-k rescale, bias, leaky_relu, residual, then split-fp16 / fp32 / clamp+quantise store
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM allocator),
// warps 2-9 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31 and half of the channels).  Persistent: each CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...; the ring of TMEM slots overlaps the epilogue with the MMAs.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

constexpr int kThreads = 320;                // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
using namespace tc;

template <int ROW_BYTES, int COUT>
struct TcCfg {
  static constexpr int A_BYTES = kTileM * ROW_BYTES;
  static constexpr int W_BYTES = COUT * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int KSLAB = ROW_BYTES / 2;       // fp16 elements per A row
  static constexpr int SLOT_COLS = 2 * COUT;        // one accumulation chain: [main | correction] columns
  static constexpr int SLOTS = 512 / SLOT_COLS > 8 ? 8 : 512 / SLOT_COLS;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + 2 * COUT * 4;
};

template <int ROW_BYTES, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
          const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
          const __grid_constant__ TcLayerParams prm, int tiles_x, int tiles_y, int num_tiles, int* error_flag) {
  using Cfg = TcCfg<ROW_BYTES, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared pointer
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                          // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* slot_full = bars + 2 * Cfg::STAGES;       // [SLOTS]   MMA -> epilogue
  uint64_t* slot_empty = slot_full + Cfg::SLOTS;      // [SLOTS]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + Cfg::SLOTS);
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);   // [2][COUT]
  static_assert((2 * Cfg::STAGES + 2 * Cfg::SLOTS) * 8 + 4 <= 256, "barrier area too small");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < Cfg::SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
  }
  for (int i = threadIdx.x; i < 2 * COUT; i += kThreads) bias_s[i] = prm.bias[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_plane = tiles_x * tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // warp-uniform control flow; one elected lane issues the copies
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int job_i = t % prm.njobs;
      int rest = t / prm.njobs;
      const int txy = rest % tiles_per_plane;
      const int p = rest / tiles_per_plane;
      const int Y0 = (txy / tiles_x) * kTileRows, X0 = (txy % tiles_x) * kTileCols;
      const int set = p < prm.n_split ? 0 : 1;
      const int nsteps = prm.jobs[job_i].nsteps;
      for (int s = 0; s < nsteps; ++s) {
        const TcStep st = prm.jobs[job_i].steps[s];
        mbar_wait(&empty_bar[stage], phase ^ 1, error_flag, 1);
        if (elect_one()) {
          uint8_t* sb = stage_base + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_5d(&map_a_hi, sb, &full_bar[stage], st.koff, X0 + st.dx, st.py, Y0 + st.dy, p);
          tma_load_5d(&map_a_lo, sb + Cfg::A_BYTES, &full_bar[stage], st.koff, X0 + st.dx, st.py, Y0 + st.dy, p);
          const int wrow = set * prm.rows_per_set + st.w_row;
          tma_load_2d(&map_w_hi, sb + 2 * Cfg::A_BYTES, &full_bar[stage], 0, wrow);
          tma_load_2d(&map_w_lo, sb + 2 * Cfg::A_BYTES + Cfg::W_BYTES, &full_bar[stage], 0, wrow);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Per k-step two MMAs:  D[slot, 0:2*COUT)    (+)= A_hi x [W_hi | W_lo]   (N = 2*COUT: hi*hi | hi*lo)
    //                       D[slot, COUT:2*COUT)  += A_lo x  W_hi             (N = COUT:   lo*hi)
    // so the slot's first COUT columns hold the main sum and the next COUT the correction terms.
    // Control flow is warp-uniform (all lanes wait on the barriers and compute the descriptors, which keeps
    // them in uniform registers); one elected lane issues the MMAs and the commits.
    constexpr uint32_t idesc_wide = make_idesc(2 * COUT);
    constexpr uint32_t idesc_narrow = make_idesc(COUT);
    const uint32_t smem_base_u32 = smem_u32(stage_base);
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int job_i = t % prm.njobs;
      const int nsteps = prm.jobs[job_i].nsteps;
      const uint32_t chain_end_mask = prm.jobs[job_i].chain_end_mask, half_mask = prm.jobs[job_i].half_mask;
      const int ks_full = prm.jobs[job_i].ks_end;
      bool chain_start = true;
      uint32_t d_tmem = 0;
      for (int s = 0; s < nsteps; ++s) {
        const int ks_end = ks_full;
        const int ks_begin = ((half_mask >> s) & 1u) ? (ks_full >> 1) : 0;
        const int chain_end = (chain_end_mask >> s) & 1u;
        if (chain_start) {
          mbar_wait(&slot_empty[slot], slot_phase ^ 1, error_flag, 2);
          d_tmem = tmem_base + slot * Cfg::SLOT_COLS;
        }
        mbar_wait(&full_bar[stage], phase, error_flag, 3);
        tc_fence_after();
        const uint32_t sb = smem_base_u32 + stage * Cfg::STAGE_BYTES;
        const uint64_t a_hi = make_smem_desc<ROW_BYTES>(sb);
        const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
        const uint64_t w_hl = a_hi + (uint64_t)((2 * Cfg::A_BYTES) >> 4);   // W_hi tile followed by the W_lo tile
        if (elect_one()) {
          if (ks_begin == 0 && ks_end == 4) {
            umma_f16(d_tmem, a_hi, w_hl, idesc_wide, chain_start ? 0u : 1u);
            umma_f16(d_tmem + COUT, a_lo, w_hl, idesc_narrow, 1u);
#pragma unroll
            for (int ks = 1; ks < 4; ++ks) {
              umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, 1u);
              umma_f16(d_tmem + COUT, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
            }
          } else {
            uint32_t accumulate = chain_start ? 0u : 1u;
            for (int ks = ks_begin; ks < ks_end; ++ks) {
              const uint64_t koff = (uint64_t)(ks * 2);     // 16 fp16 = 32 bytes, in 16-byte descriptor units
              umma_f16(d_tmem, a_hi + koff, w_hl + koff, idesc_wide, accumulate);
              umma_f16(d_tmem + COUT, a_lo + koff, w_hl + koff, idesc_narrow, 1u);
              accumulate = 1u;
            }
          }
          umma_commit(&empty_bar[stage]);     // frees the smem stage when these MMAs have read it
          if (chain_end) umma_commit(&slot_full[slot]);   // chain complete -> epilogue
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        chain_start = false;
        if (chain_end) {
          if (++slot == Cfg::SLOTS) { slot = 0; slot_phase ^= 1; }
          chain_start = true;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // 8 warps: warp w reads TMEM lanes 32*(w%4)..+31 (hardware restriction) and handles the channel half
    // (w-2)/4 of those pixels, i.e. HALF = COUT/2 channels per thread.
    constexpr int HALF = COUT / 2;
    const int lg = warp & 3;                  // TMEM lane group this warp may access
    const int hf = (warp - 2) >> 2;           // channel half
    const int m = lg * 32 + lane;             // row of the tile = pixel
    const int r = m >> 3, c = m & 7;
    const int ch0 = hf * HALF;
    int slot = 0; uint32_t slot_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int job_i = t % prm.njobs;
      int rest = t / prm.njobs;
      const int txy = rest % tiles_per_plane;
      const int p = rest / tiles_per_plane;
      const int Y = (txy / tiles_x) * kTileRows + r, X = (txy % tiles_x) * kTileCols + c;
      const int set = p < prm.n_split ? 0 : 1;
      const int nchains = prm.jobs[job_i].nchains;
      const int oy = Y * prm.out_stride + prm.jobs[job_i].out_oy, ox = X * prm.out_stride + prm.jobs[job_i].out_ox;
      const bool valid = Y < prm.Hp && X < prm.Wp && oy < prm.Ho && ox < prm.Wo;
      const float inv_scale = prm.inv_scale[set];
      const float* bs = bias_s + set * COUT + ch0;
      // sum the chains with round-to-nearest adds: acc += (main + correction)
      float acc[HALF];
#pragma unroll
      for (int i = 0; i < HALF; ++i) acc[i] = 0.0f;
      for (int ch = 0; ch < nchains; ++ch) {
        mbar_wait(&slot_full[slot], slot_phase, error_flag, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * Cfg::SLOT_COLS + ch0;
        uint32_t vm[HALF], vc[HALF];
        if (HALF == 32) { tmem_ld32_nowait(taddr, vm); tmem_ld32_nowait(taddr + COUT, vc); }
        else { tmem_ld16_nowait(taddr, vm); tmem_ld16_nowait(taddr + COUT, vc); }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[slot]);     // the slot's data is in registers now
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = __fadd_rn(acc[i], __fadd_rn(__uint_as_float(vm[i]), __uint_as_float(vc[i])));
        if (++slot == Cfg::SLOTS) { slot = 0; slot_phase ^= 1; }
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
          if (prm.clamp01) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i], 0.0f), 1.0f);
          }
          if (prm.out_mode == TC_OUT_SPLIT) {
            __align__(16) uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) split2_f32(v[i], v[i + 1], h[i / 2], l[i / 2]);
            *reinterpret_cast<uint4*>(prm.out_hi + o) = *reinterpret_cast<uint4*>(h);
            *reinterpret_cast<uint4*>(prm.out_hi + o + 8) = *reinterpret_cast<uint4*>(h + 4);
            *reinterpret_cast<uint4*>(prm.out_lo + o) = *reinterpret_cast<uint4*>(l);
            *reinterpret_cast<uint4*>(prm.out_lo + o + 8) = *reinterpret_cast<uint4*>(l + 4);
          } else if (prm.out_mode == TC_OUT_F32) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(prm.out_f32 + o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            // conv8: clip(0,1) (encoder.py:32) then np.round(e*255).astype(uint8) (encoder.py:47);
            // latent [N,Ho,Wo,96], plane-major batch p = plane*N + n, channels plane*32 + channel
            const int N = prm.P / 3;
            const int plane = p / N, n = p - plane * N;
            const size_t lo_ = (((size_t)n * prm.Ho + oy) * prm.Wo + ox) * 96 + plane * 32 + ch0 + c0;
            __align__(16) uint8_t q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = (uint8_t)rintf(__fmul_rn(v[i], 255.0f));
            *reinterpret_cast<uint4*>(prm.out_u8 + lo_) = *reinterpret_cast<uint4*>(q);
            if (prm.out_prequant) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(prm.out_prequant + lo_ + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

template <int ROW_BYTES, int COUT>
cudaError_t launch_impl(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                        const CUtensorMap& w_lo, const TcLayerParams& prm, int num_sms, int* error_flag,
                        cudaStream_t stream) {
  using Cfg = TcCfg<ROW_BYTES, COUT>;
  static bool attr_set = false;
  auto kern = k_tc_conv<ROW_BYTES, COUT>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles_x = (prm.Wp + kTileCols - 1) / kTileCols, tiles_y = (prm.Hp + kTileRows - 1) / kTileRows;
  const long long num_tiles_ll = (long long)tiles_x * tiles_y * prm.P * prm.njobs;
  if (num_tiles_ll <= 0 || num_tiles_ll > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int num_tiles = (int)num_tiles_ll;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  kern<<<grid, kThreads, Cfg::SMEM_BYTES, stream>>>(a_hi, a_lo, w_hi, w_lo, prm, tiles_x, tiles_y, num_tiles, error_flag);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_tc_conv(int row_bytes, int cout, const CUtensorMap& a_hi, const CUtensorMap& a_lo,
                           const CUtensorMap& w_hi, const CUtensorMap& w_lo, const TcLayerParams& prm, int num_sms,
                           int* error_flag, cudaStream_t stream) {
  if (row_bytes == 128 && cout == 64) return launch_impl<128, 64>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  if (row_bytes == 128 && cout == 32) return launch_impl<128, 32>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  if (row_bytes == 64 && cout == 64) return launch_impl<64, 64>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  return cudaErrorInvalidValue;
}

}  // namespace nnic
