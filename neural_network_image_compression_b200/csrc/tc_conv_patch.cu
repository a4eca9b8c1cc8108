// tcgen05 convolution kernel for every GEMM-shaped layer of the codec: conv2, conv3, conv4, conv8
// (encoder.py:11-17), dconv1, dconv5, dconv6 and the four output phases of dconv7 (decoder.py:11-16).
//
// Work item = one tile of 16 x 8 output (or, for transposed convolutions, input) pixels of one plane.  All taps of
// these layers lie in the 3x3 neighbourhood of the tile in the view the layer reads (the plain [C,W,1,H,P] view, or
// the [2C,W/2,2,H/2,P] parity view for the stride-2 convolutions), so the activation patch of a tile (18 rows x 10
// columns of pixels, 64 or 128 bytes each, hi and lo planes) is loaded ONCE per item by TMA (out-of-bounds zero fill =
// TF SAME padding) and every tap reads it through a UMMA descriptor whose start address is shifted by whole pixel
// rows and whose 8-row-group stride is the patch pitch (10 pixels); tools/probe_tc.cu (probe A) shows the tensor
// core applies the swizzle on absolute shared-memory addresses, so such descriptors read exactly the shifted rows.
// A first version streamed one [128 pixel x K] tile per tap instead and was bound by the shared-memory port (MMA
// operand reads ~125 B/clk plus TMA fills ~107 B/clk against 128 B/clk, profiles/r1_ncu_tc_v1_streaming.txt).
// Only the per-tap weight tiles [W_hi | W_lo] still stream, three taps per group.
//   * stride-2 convolutions read one patch per input-row parity (conv2: column taps are paired in one 64-wide K slab)
//     or per (row, column) parity (conv8: four patches, the column parity selects 64 of the view's 128 inner elements);
//   * dconv7 / dconv1: one item covers all four output phases of a tile, they share the patch (25 taps in total).
//
// Arithmetic: operands are fp16 hi/lo pairs; per 16-wide k-step one MMA A_hi x [W_hi | W_lo] (N = 2*COUT) and one
// A_lo x W_hi (N = COUT) accumulate into the two halves of a TMEM slot.  The tensor core truncates its fp32
// accumulator on every MMA, so a slot only takes one chain of <= 3 taps (12 k-steps); the epilogue adds main and
// correction halves and the chains with round-to-nearest fp32 adds, then bias, leaky ReLU, residual, and writes the
// next layer's hi/lo planes (or, for conv8, clip + round(x*255) into the u8 latent).
//
// Warps: 0 weight TMA, 1-2 MMA issuers (alternate chains), 4*NSPLIT epilogue warps (TMEM lane group = warp % 4,
// COUT/NSPLIT channels per thread, 256-bit global accesses), last warp patch TMA.  One persistent CTA per SM.
#include "kernels.h"
#include "tc_common.cuh"

namespace nnic {

namespace {

using namespace tc;

constexpr int kNumMma = 2;                     // MMA issuer warps (2 = alternate chains)
constexpr int kEpiWarp0 = 1 + kNumMma;        // warp 0 weight TMA, MMA issuer(s), 4*NSPLIT epilogue warps, last warp patch TMA
constexpr int PH = kTileRows + 2, PW = kTileCols + 2;        // 18 x 10 pixels
constexpr int NSETS = 2;
constexpr int GTAPS = 3;                                     // taps per accumulation chain (12 k-steps) of the 64-channel inputs
constexpr int GTAPS_K32 = 3;                                 // ... and of dconv1's 32-channel input.  NOT 6 (the same 12 k-steps), although that saves the
                                                             // epilogue four of nine chains: two issuers on alternate chains would then span 12 taps of the
                                                             // 8-slot weight ring, and a parity wait on a slot whose previous fill has not landed yet is
                                                             // satisfied by the fill before it (seen as one wrong dconv1 tile in ~1 of 400 4K decodes)
constexpr int TMEM_COLS = 512;
constexpr int F8_NT = 32;                                    // FUSE8: dconv8's 25 taps padded to the MMA N granularity

// RB = bytes of one pixel row of the patch (64 input channels -> 128, SWIZZLE_128B; 32 -> 64, SWIZZLE_64B)
// FUSE8 (dconv7 only): the epilogue hands its output tile to a second GEMM -- dconv8's tap responses, see the kernel -- which
// takes three weight-ring slots worth of shared memory (A tile hi + lo 32 KB, both dconv8 weight sets 16 KB) and one
// accumulation slot's worth of TMEM columns.
template <int RB, int NSPLIT, int COUT, bool FUSE8 = false>
struct PCfg {
  static constexpr int WSLOTS = FUSE8 ? 5 : 8;                           // weight tiles (one tap each) in flight
  static constexpr int SLOT_COLS = 2 * COUT;                             // one accumulation chain: [main | correction]
  static constexpr int SLOTS = FUSE8 ? 3 : (512 / SLOT_COLS > 8 ? 8 : 512 / SLOT_COLS);
  static constexpr int F8_COL = SLOTS * SLOT_COLS;                       // FUSE8: TMEM columns [main 32 | correction 32] of the tap responses
  static constexpr int kEpiWarps = 4 * NSPLIT;                          // NSPLIT warps per TMEM lane group, COUT/NSPLIT channels each
  static constexpr int kPatchWarp = kEpiWarp0 + kEpiWarps;
  static constexpr int kF8Warp = kPatchWarp + 1;                         // FUSE8: issuer of the response MMAs
  static constexpr int kThreads = (kPatchWarp + 1 + (FUSE8 ? 1 : 0)) * 32;
  static constexpr int KSTEPS = RB / 32;                                 // 16-element k-steps per tap
  static constexpr int PATCH_TX = PH * PW * RB;                          // bytes one patch load brings
  static constexpr int PATCH_SLOT = (PATCH_TX + 1023) / 1024 * 1024;
  static constexpr int SET_BYTES = 2 * PATCH_SLOT;                       // hi + lo
  static constexpr int W_TILE = COUT * RB;                               // one of hi / lo
  static constexpr int W_SLOT = 2 * W_TILE;                              // one tap: [W_hi | W_lo]
  static constexpr int F8_A_OFF = NSETS * SET_BYTES + WSLOTS * W_SLOT;   // FUSE8: [A_hi | A_lo], 128 pixels x 64 channels fp16 each
  static constexpr int F8_A_TILE = kTileM * 128;
  static constexpr int F8_W_OFF = F8_A_OFF + 2 * F8_A_TILE;              // FUSE8: [set][W_hi | W_lo], 32 taps x 64 channels fp16 each
  static constexpr int F8_W_SET = 2 * F8_NT * 128;
  static constexpr int BAR_OFF = FUSE8 ? F8_W_OFF + 2 * F8_W_SET : F8_A_OFF;
  static constexpr int HIST_OFF = BAR_OFF + 512 + 2 * COUT * 4;          // conv8: 256-bin histogram of the tile being quantised
  static constexpr int SMEM_BYTES = HIST_OFF + 1024 + 1024;
  static constexpr uint32_t A_SBO = PW * RB;                             // 8-row group stride of a tap view: one patch row
  static constexpr uint64_t LAYOUT = RB == 128 ? 2ull : 4ull;            // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static constexpr uint32_t W_SBO = 8 * RB;
};

// role timers (profiles/r1_tcprof_*.log) are compiled in with -DNNIC_TC_TIMERS; they cost a few hundred cycles per tile
// Development switches (NNIC_TC_DBG, kernels.h) exist only in the -DNNIC_TC_DEVELOP build (libnnic_dev.so, `make dev`), which
// tools/dbg_sweep.sh loads through NNIC_LIB; in the product library DBG() is a compile-time false and the branches vanish.
#ifdef NNIC_TC_DEVELOP
#define DBG(bit) ((prm.dbg & (bit)) != 0)
#else
#define DBG(bit) false
#endif
#ifdef NNIC_TC_TIMERS
#define TICK() ((long long)clock64())
#else
#define TICK() 0LL
#endif                         // 8-row group stride of a tap view: one patch row

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// FAST: one fp16 product per MAC (A_hi x W_hi only, activations written as one fp16 plane) -- the decoder's optional
// reduced-precision arithmetic (nnic_set_decode_precision); the lo planes are neither read nor written.
// CL = thread-block cluster size (1 or 2).  With CL = 2 the two CTAs of a cluster work on neighbouring items in lockstep
// and every weight tile is fetched from L2 once and MULTICAST into both shared memories (the CTAs alternate as the
// issuer); a ring slot is refilled when the MMAs of both CTAs have released it.
// AHI: the input activations are EXACT in one fp16 plane (dconv1 fed with the integer latent symbols 0..255, see
// nnic_api.cu decode_batch): no lo patch is loaded and the A_lo x W_hi product is not issued; everything else as in the
// split arithmetic (A_hi x [W_hi | W_lo], main + correction halves, split output).
// FUSE8 (dconv7, decoder.py:16-17,30-32): the layer's output never goes to memory.  dconv8 has ONE output channel, so it is a
// GEMM over its INPUT pixels, R[pixel, tap] = sum_ci x[pixel, ci] K8[tap, ci] (tc_dconv8.cu), and every (pixel, tap) response
// feeds exactly one output pixel.  The epilogue therefore writes each finished phase tile (128 pixels x 64 channels, hi and lo)
// into shared memory as an A operand, one thread issues the same eight MMAs k_tc_dconv8 would (A_hi x [K8_hi | K8_lo],
// A_lo x K8_hi per k-step, so the responses are bit-identical to the unfused path), and one phase later the epilogue warps move
// the 25 responses per pixel from TMEM to global memory (tile-blocked, so a warp writes whole 128-byte lines): 100 bytes per
// pixel instead of 256, read once by k_dconv8_gather.
template <int RB, int NSPLIT, int COUT, bool FAST, int CL, bool AHI, bool FUSE8, bool PIN>
__global__ void __launch_bounds__((PCfg<RB, NSPLIT, COUT, FUSE8>::kThreads), 1)
k_tc_conv_patch(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ TcPatchParams prm, int tiles_x, int tiles_y, int num_items, int* error_flag) {
  using C = PCfg<RB, NSPLIT, COUT, FUSE8>;
  static_assert(!FUSE8 || (RB == 128 && NSPLIT == 4 && COUT == 64 && !FAST && CL == 1 && !AHI), "FUSE8 is dconv7's variant");
  static_assert(!PIN || (RB == 128 && COUT == 64 && !FAST && CL == 1 && !AHI && !FUSE8), "PIN is the variant of the nine-tap layers");
  // PIN (conv3, conv4, dconv5, dconv6: one job of nine taps, one patch): seven of the nine weight tiles stay in their ring slots for
  // all items of a weight set and only taps PIN_TA and PIN_TB stream through slot 7 -- 32 KB of weight fills per item instead of
  // 144 KB.  The fills compete with the MMAs' operand reads for shared-memory bandwidth, which is what bounds these layers.
  // Issuer 0 takes chains 0 and 2 (and with them every use of the shared slot, in order), issuer 1 takes chain 1.
  constexpr int PIN_TA = 2, PIN_TB = 7, PIN_SHARED = 7;
  auto pin_slot = [](int t) { return t < PIN_TA ? t : (t < PIN_TB ? t - 1 : t - 2); };
  constexpr int WSLOTS = C::WSLOTS;
  const WaitCtx wc{error_flag, prm.wait_timeout, prm.kernel_tag};
  constexpr int kEpiWarps = C::kEpiWarps, kPatchWarp = C::kPatchWarp, kThreads = C::kThreads, SLOT_COLS = C::SLOT_COLS, SLOTS = C::SLOTS;
  constexpr int PATCH_TX = C::PATCH_TX, PATCH_SLOT = C::PATCH_SLOT, SET_BYTES = C::SET_BYTES, W_TILE = C::W_TILE, W_SLOT = C::W_SLOT,
                BAR_OFF = C::BAR_OFF, KSTEPS = C::KSTEPS;
  constexpr uint32_t A_SBO = C::A_SBO;
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
  uint64_t* f8_a_empty = slot_empty + SLOTS;   // FUSE8: the response MMAs have read the A tile
  uint64_t* f8_r_full = f8_a_empty + 1;        // FUSE8: the responses of a phase tile are in TMEM
  uint64_t* f8_a_full = f8_r_full + 1;         // FUSE8: every epilogue warp has written its rows of the A tile (and drained the previous responses)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(f8_a_full + 1);
  volatile int* tap_seen = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [kNumMma] highest tap whose weight fill an issuer has seen (RING_GUARD)
  float* bias_s = reinterpret_cast<float*>(smem + BAR_OFF + 512);
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(smem + C::HIST_OFF);
  static_assert((2 * NSETS + 2 * WSLOTS + 2 * SLOTS + 3) * 8 + 4 + 8 <= 512, "barrier area too small");
  static_assert(C::SMEM_BYTES <= 232448, "shared memory budget");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSETS; ++s) { mbar_init(&patch_full[s], 1); mbar_init(&patch_empty[s], kNumMma); }
    for (int s = 0; s < WSLOTS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CL); }
    for (int a = 0; a < SLOTS; ++a) { mbar_init(&slot_full[a], 1); mbar_init(&slot_empty[a], kEpiWarps); }
    mbar_init(f8_a_empty, 1); mbar_init(f8_r_full, 1); mbar_init(f8_a_full, kEpiWarps);
    tap_seen[0] = -1; tap_seen[1] = -1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
  }
  for (int i = threadIdx.x; i < 2 * COUT; i += kThreads) bias_s[i] = prm.bias[i];
  for (int i = threadIdx.x; i < 256; i += kThreads) hist_s[i] = 0;
  if (FUSE8) {
    // both dconv8 weight sets stay resident: [set][K8_hi rows 0-31 | K8_lo rows 32-63], 128-byte rows in the SWIZZLE_128B
    // pattern (16-byte chunk c of row r at chunk c ^ (r & 7)), as TMA would have written them
    for (int i = threadIdx.x; i < 2 * 64 * 8; i += kThreads) {
      const int set = i >> 9, r = (i >> 3) & 63, c = i & 7;
      const __half* src = (r < 32 ? prm.f8_w_hi : prm.f8_w_lo) + ((size_t)(set * F8_NT + (r & 31)) * 64 + c * 8);
      *reinterpret_cast<uint4*>(smem + C::F8_W_OFF + set * C::F8_W_SET + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(src);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync();                  // the peer's barriers exist before anything is multicast to them
  pdl_wait();                                  // everything above only touched shared memory, TMEM and the (static) weights
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_plane = tiles_x * tiles_y;
  // cluster rounds: the CTAs of a cluster take items base + rank; a CTA without an item in the last round still mirrors
  // the weight protocol (arms its barriers, releases the slots) so that its peer is never left waiting
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int base0 = (int)blockIdx.x - crank;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CL) - 1u);

  if (warp == kPatchWarp) {
    // ===================== TMA producer: activation patches =====================
    int pb = 0; uint32_t pphase = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int txy = it % tiles_per_plane, p = it / tiles_per_plane;
      const int Y0 = (txy / tiles_x) * kTileRows, X0 = (txy % tiles_x) * kTileCols;
      for (int q = 0; q < prm.npatch; ++q) {
        mbar_wait(&patch_empty[pb], pphase ^ 1, wc, 1);
        if (elect_one()) {
          uint8_t* pbuf = patch_base + pb * SET_BYTES;
          if (DBG(16) || (DBG(64) && it >= (int)blockIdx.x + NSETS * (int)gridDim.x)) {   // 64: only the first NSETS items load
            mbar_arrive(&patch_full[pb]);
          } else {
            mbar_expect_tx(&patch_full[pb], (FAST || AHI) ? PATCH_TX : 2 * PATCH_TX);
            tma_load_5d(&map_a_hi, pbuf, &patch_full[pb], prm.patch_c0[q], X0 - 1, prm.patch_py[q], Y0 - 1, p);
            if (!FAST && !AHI) tma_load_5d(&map_a_lo, pbuf + PATCH_SLOT, &patch_full[pb], prm.patch_c0[q], X0 - 1, prm.patch_py[q], Y0 - 1, p);
          }
        }
        __syncwarp();
        if (++pb == NSETS) { pb = 0; pphase ^= 1; }
      }
    }
  } else if (PIN && warp == 0) {
    // ===================== TMA producer: pinned weight tiles + the two streamed taps =====================
    int run = -1, cur_set = -1;
    uint32_t su = 0;                           // fills of the shared slot so far
    long long tw_w = 0, t_begin = TICK();
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int set = (it / tiles_per_plane) < prm.n_split ? 0 : 1;
      auto load_tap = [&](int t, int sl) {
        uint8_t* wb = w_base + sl * W_SLOT;
        mbar_expect_tx(&w_full[sl], W_SLOT);
        const int wrow = set * prm.rows_per_set + prm.jobs[0].steps[t].w_row;
        tma_load_2d(&map_w_hi, wb, &w_full[sl], 0, wrow);
        tma_load_2d(&map_w_lo, wb + W_TILE, &w_full[sl], 0, wrow);
      };
      if (set != cur_set) {                    // first item, or the Y -> CbCr boundary: (re)load the pinned tiles
        cur_set = set; ++run;
        for (int t = 0; t < 9; ++t) {
          if (t == PIN_TA || t == PIN_TB) continue;
          const int sl = pin_slot(t);
          if (run > 0) mbar_wait(&w_empty[sl], (uint32_t)(run - 1) & 1u, wc, 2);     // released after the previous set's last item
          if (elect_one()) load_tap(t, sl);
          __syncwarp();
        }
      }
      for (int k = 0; k < 2; ++k, ++su) {
        { long long t0 = TICK(); mbar_wait(&w_empty[PIN_SHARED], (su & 1u) ^ 1u, wc, 2); tw_w += TICK() - t0; }
        if (elect_one()) load_tap(k ? PIN_TB : PIN_TA, PIN_SHARED);
        __syncwarp();
      }
    }
    if (prm.dbg_buf && lane == 0) { long long* o = prm.dbg_buf + ((size_t)blockIdx.x * 4 + 0) * 8; o[0] = TICK() - t_begin; o[1] = 0; o[2] = tw_w; }
  } else if (PIN && warp < kEpiWarp0) {
    // ===================== MMA issuers, pinned weights =====================
    const int me = warp - 1;                   // 0: chains 0 and 2, 1: chain 1
    constexpr uint32_t idesc_wide = make_idesc(2 * COUT);
    constexpr uint32_t idesc_narrow = make_idesc(COUT);
    const uint32_t patch_u32 = smem_u32(patch_base), w_u32 = smem_u32(w_base);
    int pb = 0; uint32_t pphase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    int run = -1, cur_set = -1;
    long long tw_patch = 0, tw_slot = 0, tw_w = 0, t_issue = 0, t_begin = TICK();
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int set = (it / tiles_per_plane) < prm.n_split ? 0 : 1;
      if (set != cur_set) { cur_set = set; ++run; }
      const int nxt = it + (int)gridDim.x;
      const bool last_of_run = nxt >= num_items || ((nxt / tiles_per_plane) < prm.n_split ? 0 : 1) != set;
      { long long t0 = TICK(); mbar_wait(&patch_full[pb], pphase, wc, 3); tw_patch += TICK() - t0; }
      const uint32_t pset = patch_u32 + pb * SET_BYTES;
      for (int ci = 0; ci < 3; ++ci) {
        if ((ci == 1) != (me == 1)) {          // the other issuer's chain
          if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
          continue;
        }
        uint32_t a_off[GTAPS];
#pragma unroll
        for (int k = 0; k < GTAPS; ++k) a_off[k] = prm.jobs[0].steps[3 * ci + k].a_off;
        { long long t0 = TICK(); mbar_wait(&slot_empty[slot], slot_phase ^ 1, wc, 4); tw_slot += TICK() - t0; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * SLOT_COLS;
        const long long ti0 = TICK();
        if (elect_one()) {
          uint32_t accumulate = 0u;
#pragma unroll
          for (int k = 0; k < GTAPS; ++k) {
            const int t = 3 * ci + k;
            const bool stream = t == PIN_TA || t == PIN_TB;
            const int w = stream ? PIN_SHARED : pin_slot(t);
            { long long t1 = TICK(); mbar_wait(&w_full[w], stream ? (t == PIN_TB ? 1u : 0u) : ((uint32_t)run & 1u), wc, 5); tw_w += TICK() - t1; }
            tc_fence_after();
            const uint64_t a_hi = make_desc_sbo(pset + a_off[k], A_SBO, C::LAYOUT);
            const uint64_t a_lo = a_hi + (uint64_t)(PATCH_SLOT >> 4);
            const uint64_t w_hl = make_desc_sbo(w_u32 + w * W_SLOT, C::W_SBO, C::LAYOUT);   // W_hi followed by W_lo
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, accumulate);
              umma_f16(d_tmem + COUT, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
              accumulate = 1u;
            }
            if (stream || last_of_run) umma_commit(&w_empty[w]);
          }
          umma_commit(&slot_full[slot]);
        }
        __syncwarp();
        t_issue += TICK() - ti0;
        if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
      }
      if (elect_one()) umma_commit(&patch_empty[pb]);        // every MMA of this issuer has read the patch
      __syncwarp();
      if (++pb == NSETS) { pb = 0; pphase ^= 1; }
    }
    if (prm.dbg_buf && lane == 0) { long long* o = prm.dbg_buf + ((size_t)blockIdx.x * 4 + warp) * 8; o[0] = TICK() - t_begin; o[1] = tw_patch; o[2] = tw_slot; o[3] = tw_w; o[4] = t_issue; }
  } else if (warp == 0) {
    // ===================== TMA producer: weight groups =====================
    int ws = 0; uint32_t wphase = 0;
    int w_loaded = 0;
    long long tw_patch = 0, tw_w = 0, t_begin = TICK();
    for (int base = base0; base < num_items; base += gridDim.x) {
      const int it = base + crank, it_peer = base + (crank ^ 1);
      const bool active = it < num_items, peer_active = CL > 1 && it_peer < num_items;
      const int set_my = (it / tiles_per_plane) < prm.n_split ? 0 : 1;
      const int set_peer = (it_peer / tiles_per_plane) < prm.n_split ? 0 : 1;
      const int set = active ? set_my : set_peer;
      // tiles are shared unless both CTAs are active on different weight sets (Y / CbCr boundary inside the pair)
      const bool shared = CL > 1 && (!(active && peer_active) || set_my == set_peer);
      int tap = 0;
      const int nseg = prm.npatch;
      for (int seg = 0; seg < nseg; ++seg) {
      const int njobs_seg = nseg == 1 ? prm.njobs : 1;
      for (int j = 0; j < njobs_seg; ++j) {
        int sbeg = 0;
        for (int q = 0; q < seg; ++q) sbeg += prm.seg_steps[q];
        const int send = nseg == 1 ? prm.jobs[j].nsteps : sbeg + prm.seg_steps[seg];
        for (int s0 = sbeg; s0 < send; ++s0, ++tap) {          // one weight tile per tap, in the order the issuers consume them
          { long long t0 = TICK(); mbar_wait(&w_empty[ws], wphase ^ 1, wc, 2); tw_w += TICK() - t0; }
          if (elect_one()) {
            uint8_t* wb = w_base + ws * W_SLOT;
            if (DBG(4) || (DBG(32) && w_loaded >= WSLOTS)) {   // 32: only the first ring pass loads (real data, no refills)
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_expect_tx(&w_full[ws], FAST ? W_TILE : W_SLOT);
              const int wrow = set * prm.rows_per_set + prm.jobs[j].steps[s0].w_row;
              if (CL == 1) {
                tma_load_2d(&map_w_hi, wb, &w_full[ws], 0, wrow);
                if (!FAST) tma_load_2d(&map_w_lo, wb + W_TILE, &w_full[ws], 0, wrow);
              } else if (!shared || (tap & (CL - 1)) == crank) {
                const uint16_t mask = shared ? kAllCtas : (uint16_t)(1u << crank);
                tma_load_2d_multicast(&map_w_hi, wb, &w_full[ws], 0, wrow, mask);
                if (!FAST) tma_load_2d_multicast(&map_w_lo, wb + W_TILE, &w_full[ws], 0, wrow, mask);
              }
            }
          }
          __syncwarp();
          ++w_loaded;
          if (++ws == WSLOTS) { ws = 0; wphase ^= 1; }
        }
      }
      }
    }
    if (prm.dbg_buf && lane == 0) { long long* o = prm.dbg_buf + ((size_t)blockIdx.x * 4 + 0) * 8; o[0] = TICK() - t_begin; o[1] = tw_patch; o[2] = tw_w; }
  } else if (warp < kEpiWarp0) {
    // ===================== MMA issuer =====================
    const int my_parity = warp - 1;          // two issuers take alternate chains
    constexpr int GT = RB == 64 ? GTAPS_K32 : GTAPS;
    // A parity wait tells the current phase of a barrier from the previous one only, so an issuer must not wait for fill n + 1 of
    // a ring slot before fill n has been SEEN complete -- and with two issuers on alternate chains fill n may belong to the other
    // one (8-slot ring, three-tap chains: taps T and T + 1 of a chain reuse the slots of taps T - 8 and T - 7 of the other issuer's
    // last-but-one chain).  If that fill straggles while later ones land, the wait would pass on fill n - 1 and the MMAs would read a
    // stale tile.  RING_GUARD: before an issuer waits for tap g it makes sure tap g - WSLOTS was observed -- by itself (own_hist) or
    // by the other issuer (tap_seen[], written after each successful wait).  FUSE8's 5-slot ring needs it on every third tap.
#ifdef NNIC_RING_GUARD_ALL
    constexpr bool RING_GUARD = kNumMma > 1;              // before every tap, everywhere: measured +6 % on the c2 step (profiles/r2_ring_guard_ab.log)
#else
    // before every tap: FUSE8 (5-slot ring), and dconv1 (RB == 64): its taps are 2 k-steps, the shortest reuse distance of all
    // layers, and it is the launch the packed-fp32 build failed in once its epilogue stopped throttling the issuers
    constexpr bool RING_GUARD = kNumMma * GT > WSLOTS || RB == 64;
#endif
    // -DNNIC_RING_GUARD_CHAIN: the 8-slot rings check once per chain, before the TMEM-slot wait: the other issuer's observations
    // are monotone, so the largest tap index the chain needs is enough, and the check is off the per-tap critical path.  Verified
    // on the protocol model (tests/test_ring_model.py, guard "chain"); NOT yet run on a GPU (the round's GPU budget ended), hence
    // not the default -- DESIGN.md section 8.
#ifdef NNIC_RING_GUARD_CHAIN
    constexpr bool RING_GUARD_CHAIN = kNumMma > 1 && !RING_GUARD;
#else
    constexpr bool RING_GUARD_CHAIN = false;
#endif
    static_assert(GT < WSLOTS && WSLOTS <= 8, "a chain must fit the weight ring");
    int gtap = 0;                            // taps of all chains so far, in ring order
    uint32_t own_hist = 0xffffffffu;         // bit j: tap gtap - 1 - j was this issuer's (history before the first tap counts as seen)
    int chain_ctr = 0;
    long long tw_patch = 0, tw_slot = 0, tw_w = 0, t_issue = 0, t_begin = TICK();
    constexpr uint32_t idesc_wide = make_idesc(2 * COUT);
    constexpr uint32_t idesc_narrow = make_idesc(COUT);
    const uint32_t patch_u32 = smem_u32(patch_base), w_u32 = smem_u32(w_base);
    int pb = 0; uint32_t pphase = 0;
    int ws = 0; uint32_t wphase = 0;
    int slot = 0; uint32_t slot_phase = 0;
    for (int base = base0; base < num_items; base += gridDim.x) {
      const bool active = base + crank < num_items;
      for (int seg = 0; seg < prm.npatch; ++seg) {
      if (active) { long long t0 = TICK(); mbar_wait(&patch_full[pb], pphase, wc, 3); tw_patch += TICK() - t0; }
      const uint32_t pset = patch_u32 + pb * SET_BYTES;
      const int njobs_seg = prm.npatch == 1 ? prm.njobs : 1;
      for (int j = 0; j < njobs_seg; ++j) {
        int sbeg = 0;
        for (int q = 0; q < seg; ++q) sbeg += prm.seg_steps[q];
        const int send = prm.npatch == 1 ? prm.jobs[j].nsteps : sbeg + prm.seg_steps[seg];
        for (int s0 = sbeg; s0 < send; s0 += GT) {          // one chain: <= GTAPS taps into one TMEM slot
          const int ntaps = send - s0 < GT ? send - s0 : GT;
          if (kNumMma == 2 && ((chain_ctr++) & 1) != my_parity) {            // the other issuer's chain: just advance the rings
            ws += ntaps; if (ws >= WSLOTS) { ws -= WSLOTS; wphase ^= 1; }
            gtap += ntaps; own_hist <<= ntaps;
            if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
            continue;
          }
          if (RING_GUARD_CHAIN) {
            int need = -1;
#pragma unroll
            for (int k = 0; k < GT; ++k)
              if (k < ntaps && gtap + k >= WSLOTS && !((own_hist >> (WSLOTS - 1 - k)) & 1u)) need = gtap + k - WSLOTS;
            if (need >= 0 && tap_seen[my_parity ^ 1] < need) {
              const unsigned long long t0g = clock64();
              while (tap_seen[my_parity ^ 1] < need) {
                if (wc.timeout && (unsigned long long)clock64() - t0g > wc.timeout) {
                  if (wc.error_flag) atomicExch(wc.error_flag, 100 * wc.tag + 10);
                  __threadfence_system();
                  __trap();
                }
              }
            }
          }
          uint32_t a_off[GT];
#pragma unroll
          for (int k = 0; k < GT; ++k) a_off[k] = prm.jobs[j].steps[s0 + (k < ntaps ? k : 0)].a_off;
          if (active) { long long t0 = TICK(); mbar_wait(&slot_empty[slot], slot_phase ^ 1, wc, 4); tw_slot += TICK() - t0; }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * SLOT_COLS;
          const long long ti0 = TICK();
          if (elect_one()) {                     // one lane waits for each tap's weights, issues its MMAs and releases the tile
            uint32_t accumulate = 0u;
            int w = ws; uint32_t wp = wphase;
#pragma unroll
            for (int k = 0; k < GT; ++k) {
              if (k < ntaps) {
                if (RING_GUARD) {
                  const int g = gtap + k;
                  if (g >= WSLOTS && !((own_hist >> (WSLOTS - 1 - k)) & 1u)) {
                    const unsigned long long t0g = clock64();
                    while (tap_seen[my_parity ^ 1] < g - WSLOTS) {
                      if (wc.timeout && (unsigned long long)clock64() - t0g > wc.timeout) {
                        if (wc.error_flag) atomicExch(wc.error_flag, 100 * wc.tag + 10);
                        __threadfence_system();
                        __trap();
                      }
                    }
                  }
                }
                { long long t1 = TICK(); mbar_wait(&w_full[w], wp, wc, 5); tw_w += TICK() - t1; }
                if (RING_GUARD || RING_GUARD_CHAIN) tap_seen[my_parity] = gtap + k;    // volatile shared-memory accesses of one thread stay in program order
                tc_fence_after();
                if (active && !DBG(1)) {
                  const uint64_t a_hi = make_desc_sbo(pset + (a_off[k] & 0x7fffffffu), A_SBO, C::LAYOUT);
                  const uint64_t a_lo = a_hi + (uint64_t)(PATCH_SLOT >> 4);
                  const uint64_t w_hl = make_desc_sbo(w_u32 + w * W_SLOT, C::W_SBO, C::LAYOUT);   // W_hi followed by W_lo
                  const int ks_begin = (a_off[k] >> 31) ? KSTEPS / 2 : 0;
#pragma unroll
                  for (int ks = 0; ks < KSTEPS; ++ks) {
                    if (ks >= ks_begin) {
                      if (FAST) {
                        umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_narrow, accumulate);
                      } else {
                        umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, idesc_wide, accumulate);
                        if (!AHI) umma_f16(d_tmem + COUT, a_lo + 2 * ks, w_hl + 2 * ks, idesc_narrow, 1u);
                      }
                      accumulate = 1u;
                    }
                  }
                }
                if (CL > 1) umma_commit_multicast(&w_empty[w], kAllCtas); else umma_commit(&w_empty[w]);
                if (++w == WSLOTS) { w = 0; wp ^= 1; }
              }
            }
            if (active) umma_commit(&slot_full[slot]);
          }
          __syncwarp();
          ws += ntaps; if (ws >= WSLOTS) { ws -= WSLOTS; wphase ^= 1; }
          gtap += ntaps; own_hist = (own_hist << ntaps) | ((1u << ntaps) - 1u);
          t_issue += TICK() - ti0;
          if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
        }
      }
      if (active) {
        if (elect_one()) umma_commit(&patch_empty[pb]);      // every MMA of this segment has read the patch
        __syncwarp();
        if (++pb == NSETS) { pb = 0; pphase ^= 1; }
      }
      }
    }
    if (prm.dbg_buf && lane == 0) { long long* o = prm.dbg_buf + ((size_t)blockIdx.x * 4 + warp) * 8; o[0] = TICK() - t_begin; o[1] = tw_patch; o[2] = tw_slot; o[3] = tw_w; o[4] = t_issue; }
  } else if (FUSE8 && warp == C::kF8Warp) {
    // ===================== FUSE8: issuer of dconv8's response MMAs =====================
    const uint64_t a_hi = make_smem_desc<128>(smem_u32(smem + C::F8_A_OFF));
    const uint64_t a_lo = a_hi + (uint64_t)(C::F8_A_TILE >> 4);
    const uint32_t d_tmem = tmem_base + C::F8_COL;
    uint32_t k = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int set = (it / tiles_per_plane) < prm.n_split ? 0 : 1;
      const uint64_t w_hl = make_smem_desc<128>(smem_u32(smem + C::F8_W_OFF + set * C::F8_W_SET));
      for (int j = 0; j < prm.njobs; ++j, ++k) {
        if (DBG(2048)) continue;
        mbar_wait(f8_a_full, k & 1, wc, 9);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (DBG(4096)) break;
            umma_f16(d_tmem, a_hi + 2 * ks, w_hl + 2 * ks, make_idesc(2 * F8_NT), ks ? 1u : 0u);   // A_hi x [K8_hi | K8_lo]
            umma_f16(d_tmem + F8_NT, a_lo + 2 * ks, w_hl + 2 * ks, make_idesc(F8_NT), 1u);          // A_lo x K8_hi
          }
          umma_commit(f8_a_empty);
          umma_commit(f8_r_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    constexpr int HALF = COUT / NSPLIT;       // channels per thread
    const int lg = warp & 3;
    const int hf = (warp - kEpiWarp0) >> 2;
    const int ch0 = hf * HALF;
    int slot = 0; uint32_t slot_phase = 0;
    long long tw_full = 0, t_ld = 0, t_out = 0, t_begin = TICK();
    long long t_f8_wait = 0, t_f8_st = 0, t_f8_drain = 0;
    // FUSE8: number of phase tiles handed to the response GEMM so far, and where the responses of the latest one belong
    uint32_t f8_count = 0;
    size_t f8_prev_off = 0; bool f8_prev_valid = false;
    constexpr size_t f8_tap_stride = 4 * kTileM;                  // R is [item][25 taps][4 phases][128 pixels] fp32: one 128-byte line per warp and tap
    auto f8_drain = [&]() {
      // responses of phase tile f8_count - 1: this thread's pixel, taps 8 hf .. 8 hf + 7 (main + correction halves)
      mbar_wait(f8_r_full, (f8_count - 1) & 1, wc, 7);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + C::F8_COL + 8 * hf;
      uint32_t rm[8], rc[8];
      if (DBG(8) || DBG(8192)) {
#pragma unroll
        for (int t = 0; t < 8; ++t) { rm[t] = 0; rc[t] = 0; }
      } else { tmem_ld8_nowait(taddr, rm); tmem_ld8_nowait(taddr + F8_NT, rc); }
      tmem_ld_wait();
      tc_fence_before();
      if (f8_prev_valid) {
        float* dst = prm.f8_out + f8_prev_off + (size_t)(8 * hf) * f8_tap_stride;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (8 * hf + t < 25) dst[(size_t)t * f8_tap_stride] = __fadd_rn(__uint_as_float(rm[t]), __uint_as_float(rc[t]));
      }
    };
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int txy = it % tiles_per_plane;
      const int p = it / tiles_per_plane;
      const int tY0 = (txy / tiles_x) * kTileRows, tX0 = (txy % tiles_x) * kTileCols;
      const int set = p < prm.n_split ? 0 : 1;
      const float inv_scale = prm.inv_scale[set];
      const float* bs = bias_s + set * COUT + ch0;
      for (int j = 0; j < prm.njobs; ++j) {
        const int nchains = prm.jobs[j].nchains;
        float acc[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.0f;
        // every thread owns HALF consecutive channels of one pixel: 2*HALF bytes per fp16 plane, moved with 256-bit
        // accesses (one full 32-byte sector per thread and instruction)
        const int oY = tY0 + lg * 4 + (lane >> 3), oX = tX0 + (lane & 7);
        const int ooy = oY * prm.out_stride + prm.jobs[j].out_oy, oox = oX * prm.out_stride + prm.jobs[j].out_ox;
        const bool valid = oY < prm.Hp && oX < prm.Wp && ooy < prm.Ho && oox < prm.Wo && !DBG(2);
        const size_t ooff = (((size_t)p * prm.Hs + ooy) * prm.Ws + oox) * COUT + ch0;
        // residual: loaded now so that its latency hides behind the MMAs
        uint32_t res_h[HALF / 2], res_l[HALF / 2];
        const bool has_res = prm.res_hi != nullptr;
        if (has_res) {
#pragma unroll
          for (int q = 0; q < HALF / 16; ++q) {
            if (valid) {
              ld_global_v8(prm.res_hi + ooff + 16 * q, res_h + 8 * q);
              if (!FAST) ld_global_v8(prm.res_lo + ooff + 16 * q, res_l + 8 * q);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) { res_h[8 * q + e] = 0; res_l[8 * q + e] = 0; }
            }
          }
        }
        for (int ch = 0; ch < nchains; ++ch) {
          long long te0 = TICK();
          mbar_wait(&slot_full[slot], slot_phase, wc, 6);
          long long te1 = TICK(); tw_full += te1 - te0;
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + slot * SLOT_COLS + ch0;
          uint32_t vm[HALF], vc[HALF];
          if (DBG(8)) {
#pragma unroll
            for (int i = 0; i < HALF; ++i) { vm[i] = 0; vc[i] = 0; }
          } else if (HALF == 32) { tmem_ld32_nowait(taddr, vm); if (!FAST) tmem_ld32_nowait(taddr + COUT, vc); }
          else { tmem_ld16_nowait(taddr, vm); if (!FAST) tmem_ld16_nowait(taddr + COUT, vc); }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          // FUSE8: the slot of a phase's LAST chain is released only after the phase tile has been handed to the response GEMM
          // (below), so that those MMAs queue behind two chains of the main loop instead of three
          const bool f8_hold = FUSE8 && ch == nchains - 1;
          if (lane == 0 && !f8_hold) mbar_arrive(&slot_empty[slot]);
#pragma unroll
          for (int i = 0; i < HALF; ++i)
            acc[i] = __fadd_rn(acc[i], FAST ? __uint_as_float(vm[i]) : __fadd_rn(__uint_as_float(vm[i]), __uint_as_float(vc[i])));
          if (++slot == SLOTS) { slot = 0; slot_phase ^= 1; }
          t_ld += TICK() - te1;
        }
        const long long to0 = TICK();
        {
          // bias + leaky_relu in place (x*2^-k is exact, so the fused multiply-add rounds exactly like mul then add)
#pragma unroll
          for (int i = 0; i < HALF; ++i) {
            const float v = fmaf(acc[i], inv_scale, bs[i]);
            acc[i] = fmaxf(v, __fmul_rn(v, LEAKY_ALPHA));
          }
          if (has_res) {
            const __half* rh = reinterpret_cast<const __half*>(res_h);
            const __half* rl = reinterpret_cast<const __half*>(res_l);
#pragma unroll
            for (int i = 0; i < HALF; ++i)
              acc[i] = __fadd_rn(acc[i], FAST ? __half2float(rh[i]) * ACT_INV_SCALE : join_f32(rh[i], rl[i]));
          }
          if (COUT == 32 && prm.clamp01) {
#pragma unroll
            for (int i = 0; i < HALF; ++i) acc[i] = fminf(fmaxf(acc[i], 0.0f), 1.0f);
          }
          if (FUSE8 && DBG(2048)) {       // development: the layer without its fused tail
            if (lane == 0) mbar_arrive(&slot_empty[(slot + SLOTS - 1) % SLOTS]);
          } else if (FUSE8) {
            uint32_t h[HALF / 2], l[HALF / 2];
#pragma unroll
            for (int i = 0; i < HALF; i += 2) split2_f32(acc[i], acc[i + 1], h[i / 2], l[i / 2]);
            // A tile row = this thread's pixel (its TMEM lane), 16-byte chunks 2 hf and 2 hf + 1 of the 128-byte row
            const long long tf0 = TICK();
            mbar_wait(f8_a_empty, (f8_count & 1) ^ 1, wc, 8);      // the previous tile's MMAs have read it
            const long long tf1 = TICK(); t_f8_wait += tf1 - tf0;
            const int row = lg * 32 + lane;
            uint8_t* arow = smem + C::F8_A_OFF + row * 128;
#pragma unroll
            for (int q = 0; q < HALF / 8; ++q) {
              if (DBG(32768)) break;
              const int chunk = ((HALF / 8) * hf + q) ^ (row & 7);
              *reinterpret_cast<uint4*>(arow + (chunk << 4)) = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
              *reinterpret_cast<uint4*>(arow + C::F8_A_TILE + (chunk << 4)) = make_uint4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const long long tf2 = TICK(); t_f8_st += tf2 - tf1;
            if (f8_count) f8_drain();                             // the previous tile's responses leave TMEM before the next MMAs
            t_f8_drain += TICK() - tf2;
            f8_prev_off = ((size_t)it * 25 * 4 + j) * kTileM + row;
            f8_prev_valid = !DBG(2) && !DBG(16384);               // pixels beyond the plane are written too (never read)
            ++f8_count;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(f8_a_full);                             // the issuer warp takes it from here; nobody waits for it
              mbar_arrive(&slot_empty[(slot + SLOTS - 1) % SLOTS]);   // the held slot of this phase's last chain
            }
          } else if (COUT == 64 || prm.out_mode == TC_OUT_SPLIT) {
            uint32_t h[HALF / 2], l[HALF / 2];
#pragma unroll
            for (int i = 0; i < HALF; i += 2) split2_f32(acc[i], acc[i + 1], h[i / 2], l[i / 2]);
            if (valid) {
#pragma unroll
              for (int q = 0; q < HALF / 16; ++q) {
                if (DBG(1024)) {            // same bytes and store pattern, but folded into a 4 MB window: the lines are rewritten in L2 and hardly reach DRAM
                  const size_t w_ = (ooff + 16 * q) & (((size_t)1 << 20) - 1);
                  st_global_v8(prm.out_hi + w_, h + 8 * q);
                  if (!FAST) st_global_v8(prm.out_lo + w_, l + 8 * q);
                } else
                if (DBG(512)) {             // same bytes, but every store instruction of a warp writes 1 KB of CONTIGUOUS memory (wrong layout)
                  const size_t slab = ((size_t)it * prm.njobs + j) * 128 * COUT + (size_t)((warp - kEpiWarp0) * (HALF / 16) + q) * 512 + lane * 16;
                  st_global_v8(prm.out_hi + slab, h + 8 * q);
                  if (!FAST) st_global_v8(prm.out_lo + slab, l + 8 * q);
                } else
                if (DBG(128)) { st_global_2xv4(prm.out_hi + ooff + 16 * q, h + 8 * q); if (!FAST) st_global_2xv4(prm.out_lo + ooff + 16 * q, l + 8 * q); }
                else if (DBG(256)) { st_global_v8_cs(prm.out_hi + ooff + 16 * q, h + 8 * q); if (!FAST) st_global_v8_cs(prm.out_lo + ooff + 16 * q, l + 8 * q); }
                else {
                st_global_v8(prm.out_hi + ooff + 16 * q, h + 8 * q);
                if (!FAST) st_global_v8(prm.out_lo + ooff + 16 * q, l + 8 * q);
                }
              }
            }
          } else if (prm.out_mode == TC_OUT_F32) {
            if (valid) {
#pragma unroll
              for (int q = 0; q < HALF / 8; ++q) st_global_v8(prm.out_f32 + ooff + 8 * q, reinterpret_cast<const uint32_t*>(acc) + 8 * q);
            }
          } else {
            // conv8: np.round(e*255).astype(uint8) (encoder.py:47) into the latent [N,Ho,Wo,96]; plane-major batch
            // p = plane*N + n, channels plane*32 + channel
            const int N = prm.P / 3;
            const int plane = p / N, n = p - plane * N;
            const size_t lo_ = (((size_t)n * prm.Ho + ooy) * prm.Wo + oox) * 96 + plane * 32 + ch0;
            uint32_t q[HALF / 4];
#pragma unroll
            for (int i = 0; i < HALF; i += 4) {
              uint32_t w = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) w |= (uint32_t)(uint8_t)rintf(__fmul_rn(acc[i + e], 255.0f)) << (8 * e);
              q[i / 4] = w;
            }
            if (prm.hist) {
              // Histogram of the tile (tf1_13/src/training.py:62-68), counted where the symbols are produced.  All 128
              // pixels of an item belong to one (image, plane): shared-memory atomics for the non-zero symbols, zeros
              // counted per warp, then the kEpiWarps*32 = 256 epilogue threads flush one bin each.
              static_assert(COUT != 32 || kEpiWarps * 32 == 256, "one flushing thread per bin");
              uint32_t zeros = 0;
              if (valid) {
#pragma unroll
                for (int i = 0; i < HALF / 4; ++i) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const uint32_t sym = (q[i] >> (8 * e)) & 0xffu;
                    if (sym) atomicAdd(&hist_s[sym], 1u); else ++zeros;
                  }
                }
              }
              zeros = __reduce_add_sync(0xffffffffu, zeros);
              if (lane == 0 && zeros) atomicAdd(&hist_s[0], zeros);
              asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
              const int bin = (warp - kEpiWarp0) * 32 + lane;
              const uint32_t cnt = hist_s[bin];
              if (cnt) { atomicAdd(prm.hist + ((size_t)n * 3 + plane) * 256 + bin, cnt); hist_s[bin] = 0; }
              asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            }
            if (valid) {
#pragma unroll
              for (int i = 0; i < HALF / 4; i += 4)
                *reinterpret_cast<uint4*>(prm.out_u8 + lo_ + 4 * i) = make_uint4(q[i], q[i + 1], q[i + 2], q[i + 3]);
              if (prm.out_prequant) {
#pragma unroll
                for (int qq = 0; qq < HALF / 8; ++qq) st_global_v8(prm.out_prequant + lo_ + 8 * qq, reinterpret_cast<const uint32_t*>(acc) + 8 * qq);
              }
            }
          }
        }
        t_out += TICK() - to0;
      }
    }
    if (FUSE8 && f8_count && !DBG(2048)) f8_drain();
    if (prm.dbg_buf && lane == 0 && warp == kEpiWarp0) { long long* o = prm.dbg_buf + ((size_t)blockIdx.x * 4 + 3) * 8; o[0] = TICK() - t_begin; o[1] = tw_full; o[2] = t_ld; o[3] = t_out; o[4] = t_f8_wait; o[5] = t_f8_st; o[6] = t_f8_drain; }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();                  // no CTA leaves while its peer may still multicast to it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace

uint32_t tc_patch_a_offset(int dy, int dx, int row_bytes) { return (uint32_t)(((dy + 1) * PW + (dx + 1)) * row_bytes); }

template <int RB, int NSPLIT, int COUT, bool FAST = false, int CL = 1, bool AHI = false, bool FUSE8 = false, bool PIN = false>
static cudaError_t launch_patch_impl(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                     const CUtensorMap& w_lo, const TcPatchParams& prm, int num_sms, int* error_flag,
                                     cudaStream_t stream) {
  using Cfg = PCfg<RB, NSPLIT, COUT, FUSE8>;
  auto kern = k_tc_conv_patch<RB, NSPLIT, COUT, FAST, CL, AHI, FUSE8, PIN>;
  static unsigned long long attr_devices = 0;
  static int max_grid_of[64];                  // per device: CTAs that can be co-resident (persistent kernel: one wave)
  int dev = 0;
  cudaGetDevice(&dev);
  int& max_grid = max_grid_of[dev & 63];
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    max_grid = num_sms;
    if (CL > 1) {
      cudaLaunchConfig_t probe = {};
      probe.gridDim = dim3((num_sms / CL) * CL); probe.blockDim = dim3(Cfg::kThreads); probe.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute pa[1];
      pa[0].id = cudaLaunchAttributeClusterDimension; pa[0].val.clusterDim.x = CL; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
      probe.attrs = pa; probe.numAttrs = 1;
      int nclusters = 0;
      e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &probe);
      if (e != cudaSuccess) return e;
      if (nclusters < 1) return cudaErrorInvalidConfiguration;
      max_grid = nclusters * CL < num_sms ? nclusters * CL : (num_sms / CL) * CL;
    }
  }
  const int tiles_x = (prm.Wp + kTileCols - 1) / kTileCols, tiles_y = (prm.Hp + kTileRows - 1) / kTileRows;
  const long long items = (long long)tiles_x * tiles_y * prm.P;
  if (items <= 0 || items > 0x7fffffffLL) return cudaErrorInvalidValue;
  long long want = (items + CL - 1) / CL * CL;               // whole clusters
  const int grid = want < max_grid ? (int)want : max_grid;
  if (CL == 1) {
    return launch_kernel(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::SMEM_BYTES, stream, true, a_hi, a_lo, w_hi, w_lo, prm, tiles_x, tiles_y,
                         (int)items, error_flag);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::kThreads); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, w_hi, w_lo, prm, tiles_x, tiles_y, (int)items, error_flag);
}

cudaError_t launch_tc_conv_patch(int row_bytes, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                 const CUtensorMap& w_lo, const TcPatchParams& prm, int num_sms, int* error_flag,
                                 cudaStream_t stream) {
  // layers whose epilogue is the limiter (residual add, or several output phases per work item) use 16 epilogue warps
  const bool heavy_epilogue = prm.res_hi != nullptr || prm.njobs > 1;
  if (prm.fast) {
    if (prm.cout != 64 || prm.out_mode != TC_OUT_SPLIT) return cudaErrorInvalidValue;
    if (row_bytes == 128 && heavy_epilogue) return launch_patch_impl<128, 4, 64, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    if (row_bytes == 128) return launch_patch_impl<128, 2, 64, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    if (row_bytes == 64) return launch_patch_impl<64, 4, 64, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    return cudaErrorInvalidValue;
  }
  if (prm.f8_out) {                            // dconv7 handing its tiles to dconv8's response GEMM
    if (row_bytes == 128 && prm.cout == 64 && prm.njobs == 4 && !prm.res_hi && prm.f8_w_hi && prm.f8_w_lo)
      return launch_patch_impl<128, 4, 64, false, 1, false, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    return cudaErrorInvalidValue;
  }
  if (prm.a_hi_only) {                         // dconv1 on the integer latent symbols
    if (row_bytes == 64 && prm.cout == 64 && prm.out_mode == TC_OUT_SPLIT) return launch_patch_impl<64, 4, 64, false, 1, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    return cudaErrorInvalidValue;
  }
  if (prm.cluster == 2) {                      // weight tiles multicast to CTA pairs
    if (row_bytes == 128 && prm.cout == 32) return launch_patch_impl<128, 2, 32, false, 2>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    if (row_bytes == 128 && heavy_epilogue) return launch_patch_impl<128, 4, 64, false, 2>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    if (row_bytes == 128) return launch_patch_impl<128, 2, 64, false, 2>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    if (row_bytes == 64) return launch_patch_impl<64, 4, 64, false, 2>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    return cudaErrorInvalidValue;
  }
  if (row_bytes == 128 && prm.cout == 32) return launch_patch_impl<128, 2, 32>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  if (prm.pin && row_bytes == 128 && prm.cout == 64 && prm.njobs == 1 && prm.npatch == 1 && prm.jobs[0].nsteps == 9) {   // nine-tap layers
    if (heavy_epilogue) return launch_patch_impl<128, 4, 64, false, 1, false, false, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
    return launch_patch_impl<128, 2, 64, false, 1, false, false, true>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  }
  if (row_bytes == 128 && heavy_epilogue) return launch_patch_impl<128, 4, 64>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  if (row_bytes == 128) return launch_patch_impl<128, 2, 64>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  if (row_bytes == 64) return launch_patch_impl<64, 4, 64>(a_hi, a_lo, w_hi, w_lo, prm, num_sms, error_flag, stream);
  return cudaErrorInvalidValue;
}

}  // namespace nnic
