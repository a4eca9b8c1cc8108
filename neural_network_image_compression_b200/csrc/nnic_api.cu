// C ABI of libnnic.so (see include/nnic.h): handle, weight repacking, scratch arena, and the
// layer-by-layer orchestration of the encode / decode / rate paths.
#include <cuda.h>
#include <dlfcn.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nnic.h"
#include "kernels.h"

using namespace nnic;

namespace {

// ---- static layer tables (reference tf2_0/src/encoder.py:10-17, decoder.py:10-17) --------------
struct LayerSpec { const char* name; int k, s, cin, cout; bool transposed; };
const LayerSpec kEnc[5] = {{"conv1", 5, 2, 1, 32, false}, {"conv2", 5, 2, 32, 64, false}, {"conv3", 3, 1, 64, 64, false},
                           {"conv4", 3, 1, 64, 64, false}, {"conv8", 5, 2, 64, 32, false}};
const LayerSpec kDec[5] = {{"dconv1", 5, 2, 32, 64, true}, {"dconv5", 3, 1, 64, 64, true}, {"dconv6", 3, 1, 64, 64, true},
                           {"dconv7", 5, 2, 64, 64, true}, {"dconv8", 5, 2, 64, 1, true}};
// Entropynet (tf2_0/src/training.py:25-33): the three convolutions of the rate regressor; same layer types as conv2 / conv3
const LayerSpec kEnt[3] = {{"entropynet.conv1", 5, 2, 32, 64, false}, {"entropynet.conv2", 3, 1, 64, 64, false},
                           {"entropynet.conv3", 3, 1, 64, 64, false}};
inline const LayerSpec& spec_of(int set, int layer) { return set < 2 ? kEnc[layer] : kDec[layer]; }
// GEMM-shaped layer `gi` of network `net` (0 encoder: conv2, conv3, conv4, conv8; 1 decoder: dconv1, dconv5, dconv6, dconv7;
// 2 Entropynet: conv1, conv2, conv3)
inline const LayerSpec& gemm_spec(int net, int gi) { return net == 0 ? kEnc[gi + 1] : (net == 1 ? kDec[gi] : kEnt[gi]); }

thread_local std::string g_global_error = "";
thread_local const char* g_launch_ctx = "";    // layer whose launch is being enqueued (named in the error message of a failed launch)
}  // namespace
namespace nnic { int g_pdl = 1; }     // programmatic dependent launch between consecutive kernels (kernels.h launch_kernel; NNIC_PDL=0: off)
namespace {

// kernel ids reported by nnic_profile_collect (keep in sync with include/nnic.h NNIC_KERNEL_*)
enum { K_CONV1 = 0, K_CONV2, K_CONV3, K_CONV4, K_CONV8, K_QUANTISE, K_EXPAND, K_DCONV1, K_DCONV5, K_DCONV6, K_DCONV7,
       K_DCONV8, K_HIST, K_ENTROPY, K_HIST_REDUCE, K_F32_SPLIT, K_ENT_CONV, K_DENSE, K_SSIM, K_NOISE, K_COUNT };

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
};

// Per GEMM-shaped layer: the tap program and the device weight matrices of both weight sets.
struct TcLayer {
  int row_bytes = 128, cout = 64, kslab = 64;
  int njobs = 1;
  TcJob jobs[MAX_JOBS];
  int rows_per_set = 0;
  float inv_scale[2] = {1.f, 1.f};
  __half* w_hi = nullptr;      // [2][rows_per_set][kslab]
  __half* w_lo = nullptr;
  float* bias = nullptr;       // [2][cout]
  CUtensorMap map_w_hi, map_w_lo;
  bool parity_view = false;    // input read through the [2C, W/2, 2, H/2, P] view (stride-2 convs)
  int out_stride = 1;
};
struct SimtLayer {
  float* w = nullptr;          // [2][k*k][cin][cout]
  float* bias = nullptr;       // [2][cout]
};

}  // namespace

struct nnic_handle {
  int device = 0;
  int num_sms = 148;
  int arith = NNIC_ARITH_TC_SPLIT;
  std::string err;
  uint64_t launches = 0;
  int micro_batch = 0;
  bool decode_fp16 = false;         // nnic_set_decode_precision: decoder GEMM layers with one fp16 product per MAC
  uint32_t* fused_hist = nullptr;   // set around conv8's launch by encode_batch: device [nb][3][256] counts to add to
  int tc_cluster = 0;               // weight tiles multicast to CTA pairs: NNIC_TC_CLUSTER=0 never, 1 residual layers, 2 all
  bool tc_dconv8 = true;            // dconv8 on the tensor cores (NNIC_TC_DCONV8=0: FFMA kernel)
  bool int_latent = true;           // dconv1 multiplies the integer symbols and folds /255 into its epilogue (NNIC_INT_LATENT=0: x/255 split in two planes)
  bool a_hi_only = false;           // set around dconv1's launch by decode_batch when its input is the integer symbol plane
  int tc_pin = 1;                   // nine-tap layers keep seven of their weight tiles resident: 1 the residual layers conv4 / dconv6 (measured gain 6 %),
                                    // 2 conv3 / dconv5 too (no gain there), 0 never (NNIC_TC_PIN)
  bool fuse_d78 = true;             // dconv7 feeds dconv8's response GEMM on chip (NNIC_FUSE_D78=0: dconv7's output goes through memory)
  float* fuse8_out = nullptr;       // set around dconv7's launch by decode_batch: the response tensor R it writes instead of its output
  EncodeTiledFn encode_tiled = nullptr;
  unsigned long long wait_timeout = 4000000000ull;   // barrier-wait bound in SM cycles (NNIC_TC_TIMEOUT_MS, 0 = none)
  int hist_variant = 0;             // NNIC_HIST_VARIANT (development): copies * 100 + blocks per SM of k_hist
  int tc_dbg = 0;                   // development switches (NNIC_TC_DBG), read once in nnic_create
  bool tc_prof = false;             // NNIC_TC_PROF: print the in-kernel role timers of a -DNNIC_TC_TIMERS build
  long long* tc_prof_buf = nullptr; // [num_sms][4][8], per handle
  // CUtensorMaps of the activation views, keyed by (layer slot, base pointers, shape): scratch addresses repeat from call
  // to call, so a steady-state call encodes no tensor map at all (config 1 is 13 launches in ~0.25 ms: host work counts)
  struct MapKey { int slot; const void* hi; const void* lo; int P, H, W, C, Hs, Ws; bool parity;
                  bool operator==(const MapKey& o) const { return slot == o.slot && hi == o.hi && lo == o.lo && P == o.P && H == o.H && W == o.W && C == o.C && Hs == o.Hs && Ws == o.Ws && parity == o.parity; } };
  struct MapEntry { MapKey key; CUtensorMap hi, lo; };
  std::vector<MapEntry> map_cache;
  size_t map_cache_next = 0;
  uint64_t map_cache_hits = 0, map_cache_misses = 0;
  int* error_flag_host = nullptr;   // mapped pinned; written by a kernel whose barrier wait timed out
  int* error_flag_dev = nullptr;

  // host copies of the Keras-layout weights
  std::vector<float> kernel[4][5], bias[4][5];
  bool have[4][5] = {};
  bool dirty_enc = true, dirty_dec = true;

  // device weights: index 0 = encoder, 1 = decoder
  std::vector<float> w_edge[2];            // host: conv1 [2][25][32] / dconv8 [2][25][64] (passed as kernel parameters)
  std::vector<float> b_edge[2];            // host: [2][32] / [2][1]
  TcLayer tc[3][4];                        // encoder conv2,3,4,8 ; decoder dconv1,5,6,7 ; Entropynet conv1,2,3
  // conv1 on the tensor cores: [2 sets][32 channels][32 taps] fp16 hi/lo, bias [2][32]
  __half* c1_w_hi = nullptr; __half* c1_w_lo = nullptr; float* c1_bias = nullptr;
  float c1_inv_scale[2] = {1.f, 1.f};
  bool tc_conv1 = true;             // conv1 on the tensor cores (NNIC_TC_CONV1=0: FFMA kernel)
  // dconv8 on the tensor cores: [2 sets][32 taps][64 ci] fp16 hi/lo
  __half* d8_w_hi = nullptr; __half* d8_w_lo = nullptr;
  CUtensorMap d8_map_w_hi, d8_map_w_lo;
  float d8_inv_scale[2] = {1.f, 1.f};
  SimtLayer simt[3][4];
  // Entropynet (training.py:25-42): host copies of conv1..3 (Keras HWIO), dense1 [F,512], dense2 [512,1]; device dense weights
  std::vector<float> ent_kernel[5], ent_bias[5];
  bool ent_have[5] = {};
  bool ent_dirty = true;
  int ent_features = 0;                    // F = rows of dense1's kernel
  float* ent_d1_w = nullptr; float* ent_d1_b = nullptr; float* ent_d2_w = nullptr; float ent_d2_b = 0.f;

  // host-buffer calls: copies run on two internal streams and overlap the kernels of neighbouring micro-batches
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};

  // scratch arena
  DevBuf arena;
  size_t arena_used = 0;
  DevBuf rate_scratch;

  // per-kernel CUDA-event profiling (nnic_set_profiling): one (start, stop) pair per launch
  bool prof = false;
  struct ProfRec { cudaEvent_t e0, e1; int id; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  int prof_id = 0;

  // debug: tensors of the most recent encode/decode micro-batch
  struct Dbg { const __half* hi; const __half* lo; const float* f32; size_t count; size_t stored; };
  Dbg dbg[9] = {};
};

namespace {

int fail(nnic_t* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_global_error = buf;
  return code;
}
#define CK(h, call)                                                                                    \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "%s failed: %s (%s:%d%s%s)", #call, cudaGetErrorString(e__), __FILE__, __LINE__, *g_launch_ctx ? ", " : "", g_launch_ctx); \
  } while (0)
cudaEvent_t prof_event(nnic_t* h) {
  if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// launch `call` (which enqueues on stream `st`) as kernel `kid`; with profiling on, bracket it with events
#define CKL(h, kid, st, call)                                         \
  do {                                                                \
    cudaEvent_t pe0__ = nullptr, pe1__ = nullptr;                     \
    if ((h)->prof) { pe0__ = prof_event(h); pe1__ = prof_event(h); cudaEventRecord(pe0__, st); } \
    CK(h, call);                                                      \
    if ((h)->prof) { cudaEventRecord(pe1__, st); (h)->prof_recs.push_back({pe0__, pe1__, kid}); } \
    (h)->launches++;                                                  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure_buf(nnic_t* h, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  if (b.ptr) { CK(h, cudaDeviceSynchronize()); CK(h, cudaFree(b.ptr)); b.ptr = nullptr; b.bytes = 0; }
  size_t want = bytes + bytes / 8 + (1 << 20);
  CK(h, cudaMalloc(&b.ptr, want));
  b.bytes = want;
  return 0;
}
void* arena_take(nnic_t* h, size_t bytes) {
  size_t off = (h->arena_used + 1023) & ~(size_t)1023;
  h->arena_used = off + bytes;
  return static_cast<uint8_t*>(h->arena.ptr) + off;
}
inline size_t pad1k(size_t b) { return (b + 1023) & ~(size_t)1023; }

void same_pad(int in, int k, int s, int& out, int& before) {
  out = (in + s - 1) / s;
  int tot = (out - 1) * s + k - in;
  if (tot < 0) tot = 0;
  before = tot / 2;
}

// ---- tensor maps ---------------------------------------------------------------------------------
int make_map(nnic_t* h, CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, int row_bytes) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = h->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, NNIC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
// activation view [inner, X, PY, Y, P] of a split-fp16 tensor [P,H,W,C] stored with Hs x Ws pixels per plane.
// Plain view: logical extents, so everything beyond H x W is zero-filled by the TMA unit.  Parity view: the (even)
// storage extents; a padding row / column is real memory there and must hold zeros.
int make_act_map(nnic_t* h, CUtensorMap* map, const __half* base, int P, int H, int W, int C, bool parity, int kslab,
                 int row_bytes, int box_cols, int box_rows, int Hs, int Ws) {
  cuuint64_t dims[5], strides[4];
  if (!parity) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = P;
    strides[0] = (cuuint64_t)C * 2; strides[1] = (cuuint64_t)Ws * C * 2; strides[2] = (cuuint64_t)Ws * C * 2;
    strides[3] = (cuuint64_t)Hs * Ws * C * 2;
  } else {
    dims[0] = 2 * C; dims[1] = Ws / 2; dims[2] = 2; dims[3] = Hs / 2; dims[4] = P;
    strides[0] = (cuuint64_t)2 * C * 2; strides[1] = (cuuint64_t)Ws * C * 2; strides[2] = (cuuint64_t)2 * Ws * C * 2;
    strides[3] = (cuuint64_t)Hs * Ws * C * 2;
  }
  cuuint32_t box[5] = {(cuuint32_t)kslab, (cuuint32_t)box_cols, 1, (cuuint32_t)box_rows, 1};
  return make_map(h, map, const_cast<__half*>(base), 5, dims, strides, box, row_bytes);
}

// both planes' views of one activation tensor, from the handle's cache when this (tensor, shape) was seen before
int cached_act_maps(nnic_t* h, int slot, const CUtensorMap** hi, const CUtensorMap** lo, const __half* base_hi, const __half* base_lo,
                    int P, int H, int W, int C, bool parity, int kslab, int row_bytes, int box_cols, int box_rows, int Hs, int Ws) {
  const nnic_handle::MapKey key{slot, base_hi, base_lo, P, H, W, C, Hs, Ws, parity};
  for (auto& e : h->map_cache)
    if (e.key == key) { *hi = &e.hi; *lo = &e.lo; ++h->map_cache_hits; return 0; }
  ++h->map_cache_misses;
  constexpr size_t kCap = 64;
  nnic_handle::MapEntry* slot_e;
  if (h->map_cache.size() < kCap) { h->map_cache.emplace_back(); slot_e = &h->map_cache.back(); }
  else { slot_e = &h->map_cache[h->map_cache_next]; h->map_cache_next = (h->map_cache_next + 1) % kCap; }
  slot_e->key = key;
  slot_e->key.slot = -1;                         // invalid until both maps are encoded
  if (int rc = make_act_map(h, &slot_e->hi, base_hi, P, H, W, C, parity, kslab, row_bytes, box_cols, box_rows, Hs, Ws)) return rc;
  if (int rc = make_act_map(h, &slot_e->lo, base_lo, P, H, W, C, parity, kslab, row_bytes, box_cols, box_rows, Hs, Ws)) return rc;
  slot_e->key.slot = slot;
  *hi = &slot_e->hi; *lo = &slot_e->lo;
  return 0;
}

// ---- weight repacking ----------------------------------------------------------------------------
// Keras kernel element for GEMM purposes: weight from input channel ci to output channel co at tap (a,b).
inline float kval(const LayerSpec& sp, const std::vector<float>& kern, int a, int b, int ci, int co) {
  if (!sp.transposed) return kern[(((size_t)a * sp.k + b) * sp.cin + ci) * sp.cout + co];   // [kh,kw,Cin,Cout]
  return kern[(((size_t)a * sp.k + b) * sp.cout + co) * sp.cin + ci];                        // [kh,kw,Cout,Cin]
}

// Stride-2 convolution (conv2, conv8) read through the parity view: input row of output row Y and kernel row a is
// 2Y - pb + a with pb = TF's SAME pad-before of that axis (1 for an even input size, 2 for an odd one), i.e. view row
// Y + ((a - pb) >> 1) of row parity (a - pb) & 1; columns likewise.  conv2 (Cin 32) pairs the two column parities of a
// view column in one 64-wide K slab: tile (a, jj) holds kernel columns 2jj + pb + {0, 1}.
void build_s2_program(const LayerSpec& sp, int pby, int pbx, TcJob& j) {
  memset(&j, 0, sizeof j);
  if (sp.cin == 32) {
    j.nsteps = 15;
    for (int a = 0; a < 5; ++a) for (int jj = -1; jj <= 1; ++jj) {
      TcStep& s = j.steps[a * 3 + jj + 1];
      s.dy = (a - pby) >> 1; s.py = (a - pby) & 1; s.dx = jj; s.koff = 0;
      s.w_row = ((pbx == 2 ? 15 : 0) + a * 3 + jj + 1) * sp.cout;
      s.ks_begin = (pbx == 1 && jj == -1) ? 2 : 0; s.ks_end = 4;      // pad-before 1: tile jj = -1 only has the odd column
    }
  } else {
    j.nsteps = 25;
    for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) {
      TcStep& s = j.steps[a * 5 + b];
      s.dy = (a - pby) >> 1; s.py = (a - pby) & 1; s.dx = (b - pbx) >> 1; s.koff = ((b - pbx) & 1) * 64;
      s.w_row = (a * 5 + b) * sp.cout; s.ks_begin = 0; s.ks_end = 4;
    }
  }
}

void build_tc_program(const LayerSpec& sp, TcLayer& L) {
  L.cout = sp.cout;
  L.parity_view = !sp.transposed && sp.s == 2;
  L.out_stride = sp.transposed ? sp.s : 1;
  memset(L.jobs, 0, sizeof L.jobs);
  if (!sp.transposed && sp.s == 1) {            // conv3, conv4
    L.row_bytes = 128; L.kslab = 64; L.njobs = 1;
    TcJob& j = L.jobs[0]; j.nsteps = 9;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      TcStep& s = j.steps[a * 3 + b];
      s.dy = a - 1; s.dx = b - 1; s.py = 0; s.koff = 0; s.w_row = (a * 3 + b) * sp.cout; s.ks_begin = 0; s.ks_end = 4;
    }
    L.rows_per_set = 9 * sp.cout;
  } else if (!sp.transposed && sp.s == 2) {     // conv2, conv8: even input sizes (pad-before 1); other sizes per call
    L.row_bytes = 128; L.kslab = 64; L.njobs = 1;
    build_s2_program(sp, 1, 1, L.jobs[0]);
    // conv2 keeps two tile sets: rows [0, 15*64) pair the column taps for pad-before 1, rows [15*64, 30*64) for pad-before 2
    L.rows_per_set = sp.cin == 32 ? 30 * sp.cout : 25 * sp.cout;
  } else if (sp.transposed && sp.s == 1) {      // dconv5, dconv6: out[o] = sum_a x[o+1-a] K[a]
    L.row_bytes = 128; L.kslab = 64; L.njobs = 1;
    TcJob& j = L.jobs[0]; j.nsteps = 9;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      TcStep& s = j.steps[a * 3 + b];
      s.dy = 1 - a; s.dx = 1 - b; s.py = 0; s.koff = 0; s.w_row = (a * 3 + b) * sp.cout; s.ks_begin = 0; s.ks_end = 4;
    }
    L.rows_per_set = 9 * sp.cout;
  } else {                                      // dconv1 (Cin 32), dconv7 (Cin 64): four output parity phases
    L.row_bytes = sp.cin * 2; L.kslab = sp.cin; L.njobs = 4;
    for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px) {
      TcJob& j = L.jobs[py * 2 + px]; j.nsteps = 0; j.out_oy = py; j.out_ox = px;
      for (int a = (py + 1) & 1; a < 5; a += 2) for (int b = (px + 1) & 1; b < 5; b += 2) {
        TcStep& s = j.steps[j.nsteps++];
        s.dy = (py + 1 - a) / 2; s.dx = (px + 1 - b) / 2; s.py = 0; s.koff = 0; s.w_row = (a * 5 + b) * sp.cout;
        s.ks_begin = 0; s.ks_end = sp.cin / 16;
      }
    }
    L.rows_per_set = 25 * sp.cout;
  }
}

// value of the TC weight matrix of `sp` at (row, col) for Keras kernel `kern`
float tc_weight_at(const LayerSpec& sp, const std::vector<float>& kern, int row, int col) {
  const int tile = row / sp.cout, co = row % sp.cout;
  if (!sp.transposed && sp.s == 2 && sp.cin == 32) {       // conv2 paired slabs
    const int pbx = tile >= 15 ? 2 : 1, t = tile % 15;
    const int a = t / 3, jj = t % 3 - 1, px = col / 32, ci = col % 32;
    const int b = 2 * jj + pbx + px;
    if (b < 0 || b > 4) return 0.0f;
    return kval(sp, kern, a, b, ci, co);
  }
  const int a = tile / sp.k, b = tile % sp.k;
  return kval(sp, kern, a, b, col, co);
}

int upload(nnic_t* h, void** dst, const void* src, size_t bytes) {
  if (!*dst) CK(h, cudaMalloc(dst, bytes));
  CK(h, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

// Device forms of one GEMM-shaped layer for both weight sets: fp32 tap-major copies for the FFMA path and the fp16 hi/lo
// matrices + tensor maps of the tensor-core path.
int build_gemm_layer(nnic_t* h, const LayerSpec& sp, const std::vector<float>* const kerns[2], const std::vector<float>* const biases[2],
                     TcLayer& L, SimtLayer& S) {
    // --- fp32 tap-major copies for the FFMA path: [set][tap][ci][co]
    {
      const size_t per = (size_t)sp.k * sp.k * sp.cin * sp.cout;
      std::vector<float> w(2 * per), b(2 * sp.cout);
      for (int s = 0; s < 2; ++s) {
        const std::vector<float>& kern = *kerns[s];
        for (int a = 0; a < sp.k; ++a) for (int bb = 0; bb < sp.k; ++bb)
          for (int ci = 0; ci < sp.cin; ++ci) for (int co = 0; co < sp.cout; ++co)
            w[s * per + (((size_t)a * sp.k + bb) * sp.cin + ci) * sp.cout + co] = kval(sp, kern, a, bb, ci, co);
        memcpy(&b[s * sp.cout], biases[s]->data(), sp.cout * sizeof(float));
      }
      if (int rc = upload(h, (void**)&S.w, w.data(), w.size() * 4)) return rc;
      if (int rc = upload(h, (void**)&S.bias, b.data(), b.size() * 4)) return rc;
    }
    // --- tensor-core matrices: fp16 hi/lo of w * 2^kexp, [set][rows][kslab]
    build_tc_program(sp, L);
    const size_t per = (size_t)L.rows_per_set * L.kslab;
    std::vector<__half> whi(2 * per), wlo(2 * per);
    std::vector<float> b(2 * sp.cout);
    for (int s = 0; s < 2; ++s) {
      const std::vector<float>& kern = *kerns[s];
      float maxabs = 0.f;
      for (float v : kern) maxabs = fmaxf(maxabs, fabsf(v));
      int kexp = 0;
      if (maxabs > 0.f && std::isfinite(maxabs)) {
        kexp = (int)floorf(log2f(32768.0f / maxabs));
        if (kexp < -14) kexp = -14;
        if (kexp > 24) kexp = 24;
      }
      const float scale = ldexpf(1.0f, kexp);
      L.inv_scale[s] = ldexpf(1.0f, -kexp) * ACT_INV_SCALE;
      for (int r = 0; r < L.rows_per_set; ++r)
        for (int c = 0; c < L.kslab; ++c) {
          const float v = tc_weight_at(sp, kern, r, c) * scale;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          whi[s * per + (size_t)r * L.kslab + c] = hi;
          wlo[s * per + (size_t)r * L.kslab + c] = lo;
        }
      memcpy(&b[s * sp.cout], biases[s]->data(), sp.cout * sizeof(float));
    }
    if (int rc = upload(h, (void**)&L.w_hi, whi.data(), whi.size() * sizeof(__half))) return rc;
    if (int rc = upload(h, (void**)&L.w_lo, wlo.data(), wlo.size() * sizeof(__half))) return rc;
    if (int rc = upload(h, (void**)&L.bias, b.data(), b.size() * 4)) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)L.kslab, (cuuint64_t)2 * L.rows_per_set};
    cuuint64_t strides[1] = {(cuuint64_t)L.kslab * 2};
    cuuint32_t box[2] = {(cuuint32_t)L.kslab, (cuuint32_t)L.cout};
    if (int rc = make_map(h, &L.map_w_hi, L.w_hi, 2, dims, strides, box, L.row_bytes)) return rc;
    if (int rc = make_map(h, &L.map_w_lo, L.w_lo, 2, dims, strides, box, L.row_bytes)) return rc;
  return 0;
}

int finalize_weights(nnic_t* h, int net /*0 enc, 1 dec*/) {
  bool& dirty = net == 0 ? h->dirty_enc : h->dirty_dec;
  if (!dirty) return 0;
  const int set0 = net * 2;
  for (int s = 0; s < 2; ++s)
    for (int l = 0; l < 5; ++l)
      if (!h->have[set0 + s][l])
        return fail(h, NNIC_ERR_NO_WEIGHTS, "weights of set %d layer %s are not set", set0 + s, spec_of(set0, l).name);
  CK(h, cudaDeviceSynchronize());
  // edge layer (conv1 / dconv8): tap-major fp32; both Keras layouts already are [tap][..] with the unit dim dropped
  {
    const int l = net == 0 ? 0 : 4;
    const LayerSpec& sp = spec_of(set0, l);
    const int per = sp.k * sp.k * (net == 0 ? sp.cout : sp.cin);
    const int nb = sp.cout;
    std::vector<float> w(2 * per), b(2 * nb);
    for (int s = 0; s < 2; ++s) {
      memcpy(&w[s * per], h->kernel[set0 + s][l].data(), per * sizeof(float));
      memcpy(&b[s * nb], h->bias[set0 + s][l].data(), nb * sizeof(float));
    }
    h->w_edge[net] = w;
    h->b_edge[net] = b;
    if (net == 0) {
      // tensor-core form of conv1: rows = output channels, columns = K slots (25 of 32 used)
      std::vector<__half> whi(2 * 32 * 32, __float2half_rn(0.f)), wlo(2 * 32 * 32, __float2half_rn(0.f));
      for (int s = 0; s < 2; ++s) {
        float maxabs = 0.f;
        for (int i = 0; i < per; ++i) maxabs = fmaxf(maxabs, fabsf(w[s * per + i]));
        int kexp = 0;
        if (maxabs > 0.f && std::isfinite(maxabs)) {
          kexp = (int)floorf(log2f(32768.0f / maxabs));
          if (kexp < -14) kexp = -14;
          if (kexp > 24) kexp = 24;
        }
        const float scale = ldexpf(1.0f, kexp);
        h->c1_inv_scale[s] = ldexpf(1.0f, -kexp) * ACT_INV_SCALE;
        // K slot 6 * kh + kw (kw < 5); slots 5, 11, 17, 23, 29, 30, 31 stay zero (tc_conv1.cu: a kernel row = six consecutive
        // input pixels, five taps and one zero-weight slot)
        for (int t = 0; t < 25; ++t)
          for (int c = 0; c < 32; ++c) {
            const float v = w[s * per + t * 32 + c] * scale;
            const __half hi = __float2half_rn(v);
            const int kslot = 6 * (t / 5) + t % 5;
            whi[(s * 32 + c) * 32 + kslot] = hi;
            wlo[(s * 32 + c) * 32 + kslot] = __float2half_rn(v - __half2float(hi));
          }
      }
      if (int rc = upload(h, (void**)&h->c1_w_hi, whi.data(), whi.size() * sizeof(__half))) return rc;
      if (int rc = upload(h, (void**)&h->c1_w_lo, wlo.data(), wlo.size() * sizeof(__half))) return rc;
      if (int rc = upload(h, (void**)&h->c1_bias, b.data(), b.size() * 4)) return rc;
    }
    if (net == 1) {
      // tensor-core form of dconv8: rows = taps (25 of 32 used), columns = input channels
      std::vector<__half> whi(2 * 32 * 64, __float2half_rn(0.f)), wlo(2 * 32 * 64, __float2half_rn(0.f));
      for (int s = 0; s < 2; ++s) {
        float maxabs = 0.f;
        for (int i = 0; i < per; ++i) maxabs = fmaxf(maxabs, fabsf(w[s * per + i]));
        int kexp = 0;
        if (maxabs > 0.f && std::isfinite(maxabs)) {
          kexp = (int)floorf(log2f(32768.0f / maxabs));
          if (kexp < -14) kexp = -14;
          if (kexp > 24) kexp = 24;
        }
        const float scale = ldexpf(1.0f, kexp);
        h->d8_inv_scale[s] = ldexpf(1.0f, -kexp) * ACT_INV_SCALE;
        for (int t = 0; t < 25; ++t)
          for (int c = 0; c < 64; ++c) {
            const float v = w[s * per + t * 64 + c] * scale;
            const __half hi = __float2half_rn(v);
            whi[(s * 32 + t) * 64 + c] = hi;
            wlo[(s * 32 + t) * 64 + c] = __float2half_rn(v - __half2float(hi));
          }
      }
      if (int rc = upload(h, (void**)&h->d8_w_hi, whi.data(), whi.size() * sizeof(__half))) return rc;
      if (int rc = upload(h, (void**)&h->d8_w_lo, wlo.data(), wlo.size() * sizeof(__half))) return rc;
      cuuint64_t dims[2] = {64, 64};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, 32};
      if (int rc = make_map(h, &h->d8_map_w_hi, h->d8_w_hi, 2, dims, strides, box, 128)) return rc;
      if (int rc = make_map(h, &h->d8_map_w_lo, h->d8_w_lo, 2, dims, strides, box, 128)) return rc;
    }
  }
  for (int gi = 0; gi < 4; ++gi) {
    const int l = net == 0 ? gi + 1 : gi;
    const std::vector<float>* kern[2] = {&h->kernel[set0][l], &h->kernel[set0 + 1][l]};
    const std::vector<float>* bias[2] = {&h->bias[set0][l], &h->bias[set0 + 1][l]};
    if (int rc = build_gemm_layer(h, spec_of(set0, l), kern, bias, h->tc[net][gi], h->simt[net][gi])) return rc;
  }
  dirty = false;
  return 0;
}

// ---- FFMA tap programs ---------------------------------------------------------------------------
void build_simt_jobs(const LayerSpec& sp, int Hi, int Wi, SimtJobs& J) {
  memset(&J, 0, sizeof J);
  if (!sp.transposed) {
    int Ho, Wo, pt, pl;
    same_pad(Hi, sp.k, sp.s, Ho, pt);
    same_pad(Wi, sp.k, sp.s, Wo, pl);
    J.njobs = 1; J.in_stride = sp.s; J.out_stride = 1;
    SimtJob& j = J.job[0];
    for (int a = 0; a < sp.k; ++a) for (int b = 0; b < sp.k; ++b) {
      SimtTap& t = j.taps[j.ntaps++];
      t.dy = a - pt; t.dx = b - pl; t.widx = a * sp.k + b;
    }
  } else if (sp.s == 1) {
    J.njobs = 1; J.in_stride = 1; J.out_stride = 1;
    SimtJob& j = J.job[0];
    for (int a = 0; a < sp.k; ++a) for (int b = 0; b < sp.k; ++b) {
      SimtTap& t = j.taps[j.ntaps++];
      t.dy = 1 - a; t.dx = 1 - b; t.widx = a * sp.k + b;
    }
  } else {
    J.njobs = 4; J.in_stride = 1; J.out_stride = 2;
    for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px) {
      SimtJob& j = J.job[py * 2 + px];
      j.out_oy = py; j.out_ox = px;
      for (int a = (py + 1) & 1; a < 5; a += 2) for (int b = (px + 1) & 1; b < 5; b += 2) {
        SimtTap& t = j.taps[j.ntaps++];
        t.dy = (py + 1 - a) / 2; t.dx = (px + 1 - b) / 2; t.widx = a * 5 + b;
      }
    }
  }
}

struct Act {          // an activation tensor [P,H,W,C] in one of the two storage forms
  __half* hi = nullptr; __half* lo = nullptr; float* f32 = nullptr;
  int H = 0, W = 0, C = 0;
  int Hs = 0, Ws = 0;   // storage rows / columns per plane (>= H, W): even for the tensors the stride-2 parity views read
};
inline int even_up(int v) { return v + (v & 1); }

// even_storage: rows and columns are padded to even counts (the parity views [2C, W/2, 2, H/2, P] need them); the caller
// zeroes the tensor when padding was added, so that the extra row / column reads as TF's SAME zero padding
Act take_act(nnic_t* h, bool split, int P, int H, int W, int C, bool even_storage = false) {
  Act a; a.H = H; a.W = W; a.C = C;
  a.Hs = even_storage ? even_up(H) : H; a.Ws = even_storage ? even_up(W) : W;
  const size_t n = (size_t)P * a.Hs * a.Ws * C;
  if (split) { a.hi = (__half*)arena_take(h, n * 2); a.lo = (__half*)arena_take(h, n * 2); }
  else a.f32 = (float*)arena_take(h, n * 4);
  return a;
}
inline size_t act_bytes(bool split, size_t P, size_t H, size_t W, size_t C) {
  return split ? 2 * pad1k(P * H * W * C * 2) : pad1k(P * H * W * C * 4);
}
void record_dbg(nnic_t* h, int slot, const Act& a, int P) {
  h->dbg[slot] = {a.hi, a.lo, a.f32, (size_t)P * a.H * a.W * a.C, (size_t)P * a.Hs * a.Ws * a.C};
}

// A tensor-core kernel whose barrier wait timed out writes a code into mapped host memory and traps.  The flag is read
// after every host-buffer call and at the START of every later call (so callers of the asynchronous device-buffer
// path see it too, next to the sticky CUDA error the trap leaves); reading it resets it.
int check_device_error(nnic_t* h) {
  if (h->error_flag_host && *h->error_flag_host != 0) {
    const int code = *h->error_flag_host;
    *h->error_flag_host = 0;
    return fail(h, NNIC_ERR_CUDA, "a tensor-core kernel of an earlier call timed out in a barrier wait (kernel %d, wait %d); the CUDA context is unusable", code / 100 - 1, code % 100);
  }
  return 0;
}

// one GEMM-shaped layer in either arithmetic
int run_gemm_layer(nnic_t* h, int net, int gi, const Act& in, const Act& out, const Act* res, int P, int n_split,
                   int out_mode, uint8_t* out_u8, float* out_prequant, float* out_f32_planes, cudaStream_t st) {
  const LayerSpec& sp = gemm_spec(net, gi);
  const int kid = net == 0 ? K_CONV2 + gi : (net == 1 ? K_DCONV1 + gi : K_ENT_CONV);
  int Ho, Wo, Hp, Wp;
  if (!sp.transposed) { int pt; same_pad(in.H, sp.k, sp.s, Ho, pt); same_pad(in.W, sp.k, sp.s, Wo, pt); Hp = Ho; Wp = Wo; }
  else { Ho = in.H * sp.s; Wo = in.W * sp.s; Hp = in.H; Wp = in.W; }
  if (h->arith == NNIC_ARITH_SIMT_F32) {
    SimtJobs J;
    build_simt_jobs(sp, in.H, in.W, J);
    float* dst = out_mode == TC_OUT_F32 && out_f32_planes ? out_f32_planes : out.f32;
    const int clamp = (net == 0 && gi == 3) ? 1 : 0;
    CKL(h, kid, st, launch_simt_conv(sp.cin, sp.cout, in.f32, P, in.H, in.W, dst, Ho, Wo, Hp, Wp, h->simt[net][gi].w, sp.k * sp.k,
                            h->simt[net][gi].bias, res ? res->f32 : nullptr, J, n_split, clamp, st));
    return 0;
  }
  TcLayer& L = h->tc[net][gi];
  {
    const CUtensorMap *pa_hi = nullptr, *pa_lo = nullptr;
    if (L.parity_view && ((in.Hs | in.Ws) & 1)) return fail(h, NNIC_ERR_CUDA, "internal: stride-2 input without even storage");
    if (int rc = cached_act_maps(h, net * 4 + gi, &pa_hi, &pa_lo, in.hi, in.lo, P, in.H, in.W, in.C, L.parity_view, L.kslab, L.row_bytes, 10, 18, in.Hs, in.Ws)) return rc;
    TcPatchParams pp;
    memset(&pp, 0, sizeof pp);
    pp.njobs = L.njobs;
    pp.npatch = 1;
    if (L.parity_view) {
      // stride-2 convolutions read one patch per (input-row parity, inner offset) of the parity view; every patch starts one
      // view row above the tile.  conv2: steps (kernel row a, column pair jj), rows a = 0,2,4 read row parity 1, rows 1,3
      // parity 0; jj = -1 only uses the upper half of its K slab.  conv8: (row parity, column parity) = four patches.
      TcJob src;
      { int o_, pby, pbx; same_pad(in.H, sp.k, sp.s, o_, pby); same_pad(in.W, sp.k, sp.s, o_, pbx); build_s2_program(sp, pby, pbx, src); }
      TcPatchJob& dst = pp.jobs[0];
      dst.nsteps = src.nsteps; dst.nchains = 0; dst.out_oy = 0; dst.out_ox = 0;
      pp.npatch = 0;
      int n = 0;
      for (int py = 1; py >= 0; --py) for (int c0 = 64; c0 >= 0; c0 -= 64) {
        int cnt = 0;
        for (int s = 0; s < src.nsteps; ++s) {
          const TcStep& st2 = src.steps[s];
          if (st2.py != py || st2.koff != c0) continue;
          dst.steps[n].a_off = tc_patch_a_offset(st2.dy, st2.dx, L.row_bytes) | (st2.ks_begin ? 0x80000000u : 0u);
          dst.steps[n].w_row = st2.w_row;
          ++n; ++cnt;
        }
        if (!cnt) continue;
        pp.patch_py[pp.npatch] = py; pp.patch_c0[pp.npatch] = c0; pp.seg_steps[pp.npatch] = cnt;
        ++pp.npatch;
        dst.nchains += (cnt + 2) / 3;              // chains do not span patches
      }
    } else
    for (int j = 0; j < L.njobs; ++j) {
      const TcJob& src = L.jobs[j];
      TcPatchJob& dst = pp.jobs[j];
      // the patch kernel chains three taps per TMEM slot (tc_conv_patch.cu GTAPS / GTAPS_K32: the two issuers' chains must
      // fit the 8-slot weight ring together)
      const int gt = 3;
      dst.nsteps = src.nsteps; dst.nchains = (src.nsteps + gt - 1) / gt;
      dst.out_oy = src.out_oy; dst.out_ox = src.out_ox;
      for (int s = 0; s < src.nsteps; ++s) {
        dst.steps[s].a_off = tc_patch_a_offset(src.steps[s].dy, src.steps[s].dx, L.row_bytes);
        dst.steps[s].w_row = src.steps[s].w_row;
      }
    }
    pp.P = P; pp.n_split = n_split; pp.Hp = Hp; pp.Wp = Wp; pp.Ho = Ho; pp.Wo = Wo; pp.out_stride = L.out_stride;
    // storage extents of the split output (and of the residual, which shares them); fp32 / latent outputs are dense
    pp.Hs = out_mode == TC_OUT_SPLIT ? out.Hs : Ho; pp.Ws = out_mode == TC_OUT_SPLIT ? out.Ws : Wo;
    if (res && (res->Hs != out.Hs || res->Ws != out.Ws)) return fail(h, NNIC_ERR_CUDA, "internal: residual and output storage differ");
    pp.rows_per_set = L.rows_per_set;
    pp.inv_scale[0] = L.inv_scale[0]; pp.inv_scale[1] = L.inv_scale[1];
    if (h->a_hi_only) {
      // dconv1 on the integer symbols q: sum q*w is exact in the operands, and x = q/255 (decoder.py:40) enters as one factor of
      // the epilogue scale, 2^-k / 255 rounded once -- a relative 2^-24, the size of the rounding of x itself
      pp.a_hi_only = 1;
      for (int s2 = 0; s2 < 2; ++s2) pp.inv_scale[s2] = (L.inv_scale[s2] * ACT_SCALE) / 255.0f;
    }
    pp.bias = L.bias;
    pp.res_hi = res ? res->hi : nullptr; pp.res_lo = res ? res->lo : nullptr;
    pp.out_mode = out_mode;
    pp.cout = L.cout;
    pp.fast = (net == 1 && h->decode_fp16) ? 1 : 0;
    // tc_cluster: 0 never (default: no measurable gain, profiles/r1_cluster_multicast_ab.log), 1 residual layers only, 2 every layer
    pp.cluster = (!pp.fast && (h->tc_cluster == 2 || (h->tc_cluster == 1 && res))) ? 2 : 1;
    pp.pin = (h->tc_pin == 2 || (h->tc_pin == 1 && res)) ? 1 : 0;
    pp.clamp01 = (net == 0 && gi == 3) ? 1 : 0;
    pp.out_u8 = out_u8; pp.out_prequant = out_prequant;
    pp.hist = out_mode == TC_OUT_QUANT ? h->fused_hist : nullptr;
    pp.dbg = h->tc_dbg;
    pp.wait_timeout = h->wait_timeout;
    pp.kernel_tag = kid + 1;
    pp.out_hi = out.hi; pp.out_lo = out.lo;
    pp.out_f32 = out_f32_planes ? out_f32_planes : out.f32;
    if (h->fuse8_out) { pp.f8_out = h->fuse8_out; pp.f8_w_hi = h->d8_w_hi; pp.f8_w_lo = h->d8_w_lo; }
    const bool prof = h->tc_prof;
    const size_t prof_words = (size_t)h->num_sms * 4 * 8;
    if (prof) {
      if (!h->tc_prof_buf) CK(h, cudaMalloc(&h->tc_prof_buf, prof_words * sizeof(long long)));
      CK(h, cudaMemsetAsync(h->tc_prof_buf, 0, prof_words * sizeof(long long), st));
      pp.dbg_buf = h->tc_prof_buf;
    }
    g_launch_ctx = sp.name;
    CKL(h, kid, st,
        launch_tc_conv_patch(L.row_bytes, *pa_hi, *pa_lo, L.map_w_hi, L.map_w_lo, pp, h->num_sms, h->error_flag_dev, st));
    g_launch_ctx = "";
    if (prof) {
      std::vector<long long> hb(prof_words);
      cudaStreamSynchronize(st);
      cudaMemcpy(hb.data(), h->tc_prof_buf, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double a[4][8] = {};
      const int nb_ = h->num_sms;
      for (int b = 0; b < nb_; ++b) for (int r = 0; r < 4; ++r) for (int k = 0; k < 8; ++k) a[r][k] += hb[((size_t)b * 4 + r) * 8 + k] / (double)nb_;
      fprintf(stderr, "[tcprof %s] producer: total %.0f wait_patch_empty %.0f wait_w_empty %.0f | mmaA: total %.0f wait_patch %.0f wait_slot %.0f wait_w %.0f issue %.0f | "
              "mmaB: total %.0f wait_patch %.0f wait_slot %.0f wait_w %.0f issue %.0f | epi: total %.0f wait_full %.0f tmem+add %.0f out %.0f (fused dconv8: wait_a %.0f write_a %.0f drain %.0f)\n",
              sp.name, a[0][0], a[0][1], a[0][2], a[1][0], a[1][1], a[1][2], a[1][3], a[1][4], a[2][0], a[2][1], a[2][2], a[2][3], a[2][4],
              a[3][0], a[3][1], a[3][2], a[3][3], a[3][4], a[3][5], a[3][6]);
    }
    return 0;
  }
}

int pick_micro_batch(const nnic_t* h, int N, size_t pixels_per_image) {
  if (h->micro_batch > 0) {                    // caller's override, under the same plane limit as the default
    const int mb = h->micro_batch < 21845 ? h->micro_batch : 21845;
    return mb < N ? mb : N;
  }
  const size_t target = (size_t)16 << 20;   // ~16 MP of RGB pixels in flight
  size_t nb = target / (pixels_per_image ? pixels_per_image : 1);
  if (nb < 1) nb = 1;
  if (nb > 21845) nb = 21845;               // 3*nb planes must fit gridDim.y/z
  return (int)(nb < (size_t)N ? nb : (size_t)N);
}

// ---- encoder for one micro-batch -----------------------------------------------------------------
// in: rgb u8 [nb,H,W,3] (device) or f32 planes [3nb,H,W,1] (device).  Outputs are device pointers.
// hist: optional device [nb][3][256] u32, already zeroed; the symbol counts of this micro-batch are added to it
int encode_batch(nnic_t* h, const uint8_t* rgb, const float* planes, int nb, int H, int W, uint8_t* latent,
                 float* prequant, float* out_planes, uint32_t* hist, cudaStream_t st) {
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const int P = 3 * nb;
  int H1, W1, H2, W2, H3, W3, t;
  same_pad(H, 5, 2, H1, t); same_pad(W, 5, 2, W1, t);
  same_pad(H1, 5, 2, H2, t); same_pad(W1, 5, 2, W2, t);
  same_pad(H2, 5, 2, H3, t); same_pad(W2, 5, 2, W3, t);
  const size_t need = act_bytes(split, P, even_up(H1), even_up(W1), 32) + 3 * act_bytes(split, P, even_up(H2), even_up(W2), 64) +
                      (split ? 0 : act_bytes(false, P, H3, W3, 32)) + 8192;
  size_t base_used = h->arena_used;
  if (h->arena.bytes < base_used + need) return fail(h, NNIC_ERR_CUDA, "internal: arena too small (%zu < %zu)", h->arena.bytes, base_used + need);
  // conv2 and conv8 read a1 / a4 through parity views: even storage (a2, a3 share a4's layout: a2 is its residual)
  const bool even = split && h->tc_conv1;
  Act a1 = take_act(h, split, P, H1, W1, 32, even);
  Act a2 = take_act(h, split, P, H2, W2, 64, even);
  Act a3 = take_act(h, split, P, H2, W2, 64, even);
  Act a4 = take_act(h, split, P, H2, W2, 64, even);
  if (split && !even && ((H1 | W1 | H2 | W2) & 1))
    return fail(h, NNIC_ERR_SHAPE, "NNIC_TC_CONV1=0 (development) only supports sizes that are multiples of 8");
  if (a1.Hs != H1 || a1.Ws != W1) {             // the padding row / column must read as zeros
    CK(h, cudaMemsetAsync(a1.hi, 0, (size_t)P * a1.Hs * a1.Ws * 32 * 2, st));
    CK(h, cudaMemsetAsync(a1.lo, 0, (size_t)P * a1.Hs * a1.Ws * 32 * 2, st));
  }
  if (a4.Hs != H2 || a4.Ws != W2) {
    CK(h, cudaMemsetAsync(a4.hi, 0, (size_t)P * a4.Hs * a4.Ws * 64 * 2, st));
    CK(h, cudaMemsetAsync(a4.lo, 0, (size_t)P * a4.Hs * a4.Ws * 64 * 2, st));
  }
  Act a5; a5.H = H3; a5.W = W3; a5.C = 32;
  if (split && h->tc_conv1) {
    TcConv1Params cp;
    memset(&cp, 0, sizeof cp);
    cp.rgb = rgb; cp.planes = planes; cp.N = nb; cp.H = H; cp.W = W; cp.Ho = H1; cp.Wo = W1;
    int tmp;
    same_pad(H, 5, 2, tmp, cp.pad_t); same_pad(W, 5, 2, tmp, cp.pad_l);
    cp.w_hi = h->c1_w_hi; cp.w_lo = h->c1_w_lo; cp.bias = h->c1_bias;
    cp.inv_scale[0] = h->c1_inv_scale[0]; cp.inv_scale[1] = h->c1_inv_scale[1];
    cp.cc = colour_consts();
    cp.wait_timeout = h->wait_timeout;
    cp.out_hi = a1.hi; cp.out_lo = a1.lo; cp.Hs = a1.Hs; cp.Ws = a1.Ws;
    CKL(h, K_CONV1, st, launch_tc_conv1(cp, h->num_sms, h->error_flag_dev, st));
  } else {
    CKL(h, K_CONV1, st, launch_conv1(rgb, planes, nb, H, W, h->w_edge[0].data(), h->b_edge[0].data(), a1.hi, a1.lo, a1.f32, st));
  }
  if (int rc = run_gemm_layer(h, 0, 0, a1, a2, nullptr, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return rc;
  if (int rc = run_gemm_layer(h, 0, 1, a2, a3, nullptr, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return rc;
  if (int rc = run_gemm_layer(h, 0, 2, a3, a4, &a2, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return rc;
  if (split) {
    if (out_planes) {
      if (int rc = run_gemm_layer(h, 0, 3, a4, a5, nullptr, P, nb, TC_OUT_F32, nullptr, nullptr, out_planes, st)) return rc;
    } else {
      h->fused_hist = hist;                      // conv8's epilogue counts the symbols it writes
      const int rc = run_gemm_layer(h, 0, 3, a4, a5, nullptr, P, nb, TC_OUT_QUANT, latent, prequant, nullptr, st);
      h->fused_hist = nullptr;
      if (rc) return rc;
    }
  } else {
    if (out_planes) {
      if (int rc = run_gemm_layer(h, 0, 3, a4, a5, nullptr, P, nb, TC_OUT_F32, nullptr, nullptr, out_planes, st)) return rc;
    } else {
      a5 = take_act(h, false, P, H3, W3, 32);
      if (int rc = run_gemm_layer(h, 0, 3, a4, a5, nullptr, P, nb, TC_OUT_F32, nullptr, nullptr, nullptr, st)) return rc;
      CKL(h, K_QUANTISE, st, launch_quantise(a5.f32, nb, H3, W3, latent, prequant, st));
      if (hist) CKL(h, K_HIST, st, launch_hist(latent, nb, (size_t)H3 * W3, hist, h->num_sms, h->hist_variant, st));
    }
  }
  record_dbg(h, 0, a1, P); record_dbg(h, 1, a2, P); record_dbg(h, 2, a3, P); record_dbg(h, 3, a4, P);
  h->arena_used = base_used;
  return 0;
}

// dconv7's fused output: 16 x 8-pixel tiles of its INPUT grid (2lh x 2lw), 25 taps x 4 phases x 128 pixels fp32 per tile
size_t d78_response_bytes(size_t P, int lh, int lw) {
  const size_t tiles = (size_t)((2 * lh + 15) / 16) * ((2 * lw + 7) / 8);
  return pad1k(P * tiles * 12800 * 4);
}
size_t dec_act_need(bool split, size_t P, int lh, int lw) {
  const size_t d4 = act_bytes(split, P, 4 * lh, 4 * lw, 64), resp = d78_response_bytes(P, lh, lw);
  return act_bytes(split, P, lh, lw, 32) + 3 * act_bytes(split, P, 2 * lh, 2 * lw, 64) + (d4 > resp ? d4 : resp) + 16384;
}

// ---- decoder for one micro-batch -----------------------------------------------------------------
int decode_batch(nnic_t* h, const uint8_t* latent, const float* planes, int nb, int lh, int lw, uint8_t* rgb,
                 float* prequant, float* out_planes, cudaStream_t st) {
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const int P = 3 * nb;
  const size_t need = dec_act_need(split, P, lh, lw) - 8192;
  size_t base_used = h->arena_used;
  if (h->arena.bytes < base_used + need) return fail(h, NNIC_ERR_CUDA, "internal: arena too small (%zu < %zu)", h->arena.bytes, base_used + need);
  Act d0 = take_act(h, split, P, lh, lw, 32);
  Act d1 = take_act(h, split, P, 2 * lh, 2 * lw, 64);
  Act d2 = take_act(h, split, P, 2 * lh, 2 * lw, 64);
  Act d3 = take_act(h, split, P, 2 * lh, 2 * lw, 64);
  // dconv7 + dconv8: dconv7 hands its tiles to dconv8's response GEMM on chip and writes R, 25 fp32 tap responses per output
  // pixel (100 bytes instead of 256) in tile-blocked order [P][tile][25][4 phases][128 pixels], which k_dconv8_gather turns into RGB
  const bool fuse78 = split && h->tc_dconv8 && h->fuse_d78 && !h->decode_fp16 && h->tc_cluster != 2;
  Act d4; d4.H = 4 * lh; d4.W = 4 * lw; d4.C = 64; d4.Hs = d4.H; d4.Ws = d4.W;
  float* resp = nullptr;
  if (fuse78) resp = (float*)arena_take(h, d78_response_bytes(P, lh, lw));
  else d4 = take_act(h, split, P, 4 * lh, 4 * lw, 64);
  // u8 latent into the tensor-core decoder: the symbols go in as exact fp16 integers (one plane, one product less per MAC)
  const bool int_latent = latent && split && !h->decode_fp16 && h->int_latent;
  if (latent) {
    CKL(h, K_EXPAND, st, launch_latent_expand(latent, nb, lh, lw, d0.hi, d0.lo, d0.f32, int_latent, st));
  } else if (split) {
    CKL(h, K_F32_SPLIT, st, launch_f32_to_split(planes, (size_t)P * lh * lw * 32, d0.hi, d0.lo, st));
  } else {
    d0.f32 = const_cast<float*>(planes);
  }
  h->a_hi_only = int_latent;
  const int rc_d1 = run_gemm_layer(h, 1, 0, d0, d1, nullptr, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st);
  h->a_hi_only = false;
  if (rc_d1) return rc_d1;
  if (int rc = run_gemm_layer(h, 1, 1, d1, d2, nullptr, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return rc;
  if (int rc = run_gemm_layer(h, 1, 2, d2, d3, &d1, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return rc;
  h->fuse8_out = resp;
  const int rc_d7 = run_gemm_layer(h, 1, 3, d3, d4, nullptr, P, nb, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st);
  h->fuse8_out = nullptr;
  if (rc_d7) return rc_d7;
  if (fuse78) {
    TcDconv8Params dp;
    memset(&dp, 0, sizeof dp);
    dp.N = nb; dp.Hi = 4 * lh; dp.Wi = 4 * lw;
    dp.inv_scale[0] = h->d8_inv_scale[0]; dp.inv_scale[1] = h->d8_inv_scale[1];
    dp.bias[0] = h->b_edge[1][0]; dp.bias[1] = h->b_edge[1][1];
    dp.cc = colour_consts();
    dp.rgb = rgb; dp.prequant = prequant; dp.planes_out = out_planes;
    CKL(h, K_DCONV8, st, launch_dconv8_gather(resp, dp, 2 * lh, 2 * lw, st));
  } else if (split && h->tc_dconv8) {
    const CUtensorMap *ma_hi = nullptr, *ma_lo = nullptr;
    if (int rc = cached_act_maps(h, 15, &ma_hi, &ma_lo, d4.hi, d4.lo, P, 4 * lh, 4 * lw, 64, false, 64, 128, 8, 16, d4.Hs, d4.Ws)) return rc;
    TcDconv8Params dp;
    memset(&dp, 0, sizeof dp);
    dp.N = nb; dp.Hi = 4 * lh; dp.Wi = 4 * lw;
    dp.fast = h->decode_fp16 ? 1 : 0;
    dp.wait_timeout = h->wait_timeout;
    dp.inv_scale[0] = h->d8_inv_scale[0]; dp.inv_scale[1] = h->d8_inv_scale[1];
    dp.bias[0] = h->b_edge[1][0]; dp.bias[1] = h->b_edge[1][1];
    dp.cc = colour_consts();
    dp.rgb = rgb; dp.prequant = prequant; dp.planes_out = out_planes;
    CKL(h, K_DCONV8, st, launch_tc_dconv8(*ma_hi, *ma_lo, h->d8_map_w_hi, h->d8_map_w_lo, dp, h->num_sms, h->error_flag_dev, st));
  } else {
    CKL(h, K_DCONV8, st, launch_dconv8(d4.hi, d4.lo, d4.f32, nb, 4 * lh, 4 * lw, h->w_edge[1].data(), h->b_edge[1].data(), rgb, prequant, out_planes, st));
  }
  record_dbg(h, 4, d0, P); record_dbg(h, 5, d1, P); record_dbg(h, 6, d2, P); record_dbg(h, 7, d3, P);
  if (fuse78) h->dbg[8] = {}; else record_dbg(h, 8, d4, P);       // fused: dconv7's output exists on chip only
  h->arena_used = base_used;
  return 0;
}

// after a failed host-buffer call: wait for everything the call enqueued (errors ignored: the call already failed)
void drain_streams(nnic_t* h, cudaStream_t st) {
  cudaStreamSynchronize(h->h2d_stream);
  cudaStreamSynchronize(st);
  cudaStreamSynchronize(h->d2h_stream);
  cudaGetLastError();
}

size_t enc_act_need(bool split, size_t P, int H, int W) {
  int H1, W1, H2, W2, H3, W3, t;
  same_pad(H, 5, 2, H1, t); same_pad(W, 5, 2, W1, t);
  same_pad(H1, 5, 2, H2, t); same_pad(W1, 5, 2, W2, t);
  same_pad(H2, 5, 2, H3, t); same_pad(W2, 5, 2, W3, t);
  return act_bytes(split, P, even_up(H1), even_up(W1), 32) + 3 * act_bytes(split, P, even_up(H2), even_up(W2), 64) +
         act_bytes(false, P, H3, W3, 32) + 16384;
}


// ---- Entropynet (tf2_0/src/training.py:25-42) -------------------------------------------------------------------
int finalize_entropynet(nnic_t* h) {
  if (!h->ent_dirty) return 0;
  static const char* names[5] = {"conv1", "conv2", "conv3", "dense1", "dense2"};
  for (int l = 0; l < 5; ++l)
    if (!h->ent_have[l]) return fail(h, NNIC_ERR_NO_WEIGHTS, "Entropynet weights of layer %s are not set", names[l]);
  CK(h, cudaDeviceSynchronize());
  for (int gi = 0; gi < 3; ++gi) {               // one network for every plane: both weight "sets" hold the same kernel
    const std::vector<float>* kern[2] = {&h->ent_kernel[gi], &h->ent_kernel[gi]};
    const std::vector<float>* bias[2] = {&h->ent_bias[gi], &h->ent_bias[gi]};
    if (int rc = build_gemm_layer(h, kEnt[gi], kern, bias, h->tc[2][gi], h->simt[2][gi])) return rc;
  }
  if (h->ent_d1_w) { cudaFree(h->ent_d1_w); h->ent_d1_w = nullptr; }       // its size follows the feature count
  if (int rc = upload(h, (void**)&h->ent_d1_w, h->ent_kernel[3].data(), h->ent_kernel[3].size() * 4)) return rc;
  if (int rc = upload(h, (void**)&h->ent_d1_b, h->ent_bias[3].data(), 512 * 4)) return rc;
  if (int rc = upload(h, (void**)&h->ent_d2_w, h->ent_kernel[4].data(), 512 * 4)) return rc;
  h->ent_d2_b = h->ent_bias[4][0];
  h->ent_dirty = false;
  return 0;
}


// NNIC_MEM_DEVICE buffers must be device (or managed) memory of the handle's GPU: a host pointer or another GPU's memory
// would fault inside a kernel, so it is rejected before anything is enqueued.  NULL (optional outputs) passes.
int check_device_ptrs(nnic_t* h, int mem_kind, std::initializer_list<std::pair<const char*, const void*>> ptrs) {
  if (mem_kind != NNIC_MEM_DEVICE) return 0;
  for (const auto& np : ptrs) {
    if (!np.second) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, np.second) != cudaSuccess) { cudaGetLastError(); return fail(h, NNIC_ERR_INVALID_ARG, "%s: not a CUDA pointer", np.first); }
    if (at.type == cudaMemoryTypeManaged) continue;
    if (at.type != cudaMemoryTypeDevice || at.device != h->device)
      return fail(h, NNIC_ERR_INVALID_ARG, "%s: NNIC_MEM_DEVICE buffers must be device memory of GPU %d", np.first, h->device);
  }
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* nnic_version(void) { return "nnic-b200 0.1 (sm_100a)"; }

const char* nnic_last_error(const nnic_t* h) { return h ? h->err.c_str() : g_global_error.c_str(); }

int nnic_create(int device, nnic_t** out) {
  if (!out) return fail(nullptr, NNIC_ERR_INVALID_ARG, "nnic_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(nullptr, NNIC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, NNIC_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, NNIC_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, NNIC_ERR_NO_DEVICE, "device %d is sm_%d%d; libnnic is built for sm_100a only", device, prop.major, prop.minor);
  nnic_t* h = new nnic_t();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  DeviceGuard g(device);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    delete h;
    return fail(nullptr, NNIC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  }
  h->encode_tiled = (EncodeTiledFn)fn;
  if (const char* env = getenv("NNIC_TC_DCONV8")) h->tc_dconv8 = atoi(env) != 0;
  if (const char* env = getenv("NNIC_TC_CONV1")) h->tc_conv1 = atoi(env) != 0;
  if (const char* env = getenv("NNIC_INT_LATENT")) h->int_latent = atoi(env) != 0;
  if (const char* env = getenv("NNIC_FUSE_D78")) h->fuse_d78 = atoi(env) != 0;
  if (const char* env = getenv("NNIC_TC_PIN")) h->tc_pin = atoi(env);
  if (const char* env = getenv("NNIC_PDL")) nnic::g_pdl = atoi(env) != 0;    // process-wide
  if (const char* env = getenv("NNIC_TC_CLUSTER")) h->tc_cluster = atoi(env);
  if (const char* env = getenv("NNIC_TC_DBG")) h->tc_dbg = atoi(env);
  if (const char* env = getenv("NNIC_HIST_VARIANT")) h->hist_variant = atoi(env);
  h->tc_prof = getenv("NNIC_TC_PROF") != nullptr;
  if (const char* env = getenv("NNIC_TC_TIMEOUT_MS")) {          // 0: never trap (profilers, debuggers, MPS)
    const double ms_ = atof(env);
    h->wait_timeout = ms_ <= 0.0 ? 0ull : (unsigned long long)(ms_ * 2.0e6);   // SM cycles at ~2 GHz
  }
  h->map_cache.reserve(64);
  e = cudaHostAlloc((void**)&h->error_flag_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *h->error_flag_host = 0; e = cudaHostGetDevicePointer((void**)&h->error_flag_dev, h->error_flag_host, 0); }
  if (e != cudaSuccess) {
    // no mapped host memory (e.g. under CUDA_ENABLE_COREDUMP_ON_EXCEPTION): run without the timeout report; a timed-out wait still traps
    if (h->error_flag_host) cudaFreeHost(h->error_flag_host);
    h->error_flag_host = nullptr; h->error_flag_dev = nullptr;
    cudaGetLastError();
  }
  bool ok = cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i)
    ok = cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { nnic_destroy(h); return fail(nullptr, NNIC_ERR_CUDA, "stream / event creation failed"); }
  *out = h;
  return NNIC_OK;
}

void nnic_destroy(nnic_t* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  for (int n = 0; n < 3; ++n) {
    for (int i = 0; i < 4; ++i) {
      cudaFree(h->tc[n][i].w_hi); cudaFree(h->tc[n][i].w_lo); cudaFree(h->tc[n][i].bias);
      cudaFree(h->simt[n][i].w); cudaFree(h->simt[n][i].bias);
    }
  }
  cudaFree(h->d8_w_hi); cudaFree(h->d8_w_lo);
  cudaFree(h->ent_d1_w); cudaFree(h->ent_d1_b); cudaFree(h->ent_d2_w);
  cudaFree(h->c1_w_hi); cudaFree(h->c1_w_lo); cudaFree(h->c1_bias);
  cudaFree(h->arena.ptr); cudaFree(h->rate_scratch.ptr); cudaFree(h->tc_prof_buf);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  for (int i = 0; i < 2; ++i) { if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]); if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]); if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]); }
  if (h->error_flag_host) cudaFreeHost(h->error_flag_host);
  for (auto& r : h->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : h->prof_pool) cudaEventDestroy(e);
  delete h;
}

int nnic_set_arith(nnic_t* h, int arith) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (arith != NNIC_ARITH_TC_SPLIT && arith != NNIC_ARITH_SIMT_F32) return fail(h, NNIC_ERR_INVALID_ARG, "unknown arithmetic %d", arith);
  h->arith = arith;
  return NNIC_OK;
}
int nnic_set_decode_precision(nnic_t* h, int precision) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (precision != NNIC_DECODE_SPLIT && precision != NNIC_DECODE_FP16)
    return fail(h, NNIC_ERR_INVALID_ARG, "nnic_set_decode_precision: unknown precision %d", precision);
  h->decode_fp16 = precision == NNIC_DECODE_FP16;
  return NNIC_OK;
}
int nnic_get_decode_precision(const nnic_t* h) { return h && h->decode_fp16 ? NNIC_DECODE_FP16 : NNIC_DECODE_SPLIT; }

int nnic_get_arith(const nnic_t* h) { return h ? h->arith : NNIC_ERR_INVALID_ARG; }
uint64_t nnic_launch_count(const nnic_t* h) { return h ? h->launches : 0; }
int nnic_set_micro_batch(nnic_t* h, int n) { if (!h || n < 0) return NNIC_ERR_INVALID_ARG; h->micro_batch = n; return NNIC_OK; }
uint64_t nnic_tensor_map_encodes(const nnic_t* h) { return h ? 2 * h->map_cache_misses : 0; }
size_t nnic_scratch_bytes(const nnic_t* h) { return h ? h->arena.bytes + h->rate_scratch.bytes : 0; }

void nnic_colour_constants(float* k9, float* kinv9, float* off3) {
  const ColourConsts& c = colour_consts();
  if (k9) memcpy(k9, c.k, 9 * sizeof(float));
  if (kinv9) memcpy(kinv9, c.kinv, 9 * sizeof(float));
  if (off3) memcpy(off3, c.off, 3 * sizeof(float));
}

int nnic_set_weights(nnic_t* h, int set, int layer, const float* kernel, const float* bias) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (set < 0 || set > 3 || layer < 0 || layer >= NNIC_LAYERS_PER_NET) return fail(h, NNIC_ERR_INVALID_ARG, "bad set/layer %d/%d", set, layer);
  if (!kernel || !bias) return fail(h, NNIC_ERR_INVALID_ARG, "kernel/bias is NULL");
  const LayerSpec& sp = spec_of(set, layer);
  const size_t nk = (size_t)sp.k * sp.k * sp.cin * sp.cout;
  h->kernel[set][layer].assign(kernel, kernel + nk);
  h->bias[set][layer].assign(bias, bias + sp.cout);
  h->have[set][layer] = true;
  if (set < 2) h->dirty_enc = true; else h->dirty_dec = true;
  return NNIC_OK;
}

// d_hist: optional DEVICE buffer [N][3][256] u32 (zeroed by the caller on `stream`) that receives the symbol counts
static int encode_impl(nnic_t* h, const uint8_t* rgb, int N, int H, int W, uint8_t* latent, float* prequant, uint32_t* d_hist,
                       int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!rgb || !latent) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_encode: NULL buffer");
  if (N <= 0 || H <= 0 || W <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_encode: non-positive shape %dx%dx%d", N, H, W);
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"rgb", rgb}, {"latent", latent}, {"prequant", prequant}})) return rc;
  if (int rc = check_device_error(h)) return rc;
  if (int rc = finalize_weights(h, 0)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const int lh = (H + 7) / 8, lw = (W + 7) / 8;
  const size_t img_px = (size_t)H * W, lat_px = (size_t)lh * lw;
  const bool host = mem_kind == NNIC_MEM_HOST;
  int mb = pick_micro_batch(h, N, img_px);
  if (host && h->micro_batch == 0 && N >= 2) {      // at least ~4 micro-batches so copies overlap the kernels
    const int quarter = (N + 3) / 4;
    if (quarter < mb) mb = quarter;
  }
  size_t need = enc_act_need(split, 3 * (size_t)mb, H, W);
  if (host) need += 2 * (pad1k(mb * img_px * 3) + pad1k(mb * lat_px * 96) + (prequant ? pad1k(mb * lat_px * 96 * 4) : 0)) + 8192;
  if (int rc = ensure_buf(h, h->arena, need)) return rc;
  h->arena_used = 0;
  uint8_t *d_rgb[2] = {nullptr, nullptr}, *d_lat[2] = {nullptr, nullptr}; float* d_pre[2] = {nullptr, nullptr};
  if (host) {
    for (int s2 = 0; s2 < 2; ++s2) {
      d_rgb[s2] = (uint8_t*)arena_take(h, mb * img_px * 3);
      d_lat[s2] = (uint8_t*)arena_take(h, mb * lat_px * 96);
      if (prequant) d_pre[s2] = (float*)arena_take(h, mb * lat_px * 96 * 4);
    }
  }
  auto run = [&]() -> int {
    int idx = 0;
    for (int i0 = 0; i0 < N; i0 += mb, ++idx) {
      const int nb = (N - i0) < mb ? (N - i0) : mb;
      const uint8_t* src = rgb + (size_t)i0 * img_px * 3;
      uint8_t* dst = latent + (size_t)i0 * lat_px * 96;
      float* pre = prequant ? prequant + (size_t)i0 * lat_px * 96 : nullptr;
      if (host) {
        const int s2 = idx & 1;
        if (idx >= 2) CK(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_out[s2], 0));   // staging set s2 is drained
        CK(h, cudaMemcpyAsync(d_rgb[s2], src, nb * img_px * 3, cudaMemcpyHostToDevice, h->h2d_stream));
        CK(h, cudaEventRecord(h->ev_in[s2], h->h2d_stream));
        CK(h, cudaStreamWaitEvent(st, h->ev_in[s2], 0));
        if (int rc = encode_batch(h, d_rgb[s2], nullptr, nb, H, W, d_lat[s2], d_pre[s2], nullptr, d_hist ? d_hist + (size_t)i0 * 768 : nullptr, st)) return rc;
        CK(h, cudaEventRecord(h->ev_comp[s2], st));
        CK(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_comp[s2], 0));
        CK(h, cudaMemcpyAsync(dst, d_lat[s2], nb * lat_px * 96, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (pre) CK(h, cudaMemcpyAsync(pre, d_pre[s2], nb * lat_px * 96 * 4, cudaMemcpyDeviceToHost, h->d2h_stream));
        CK(h, cudaEventRecord(h->ev_out[s2], h->d2h_stream));
      } else {
        if (int rc = encode_batch(h, src, nullptr, nb, H, W, dst, pre, nullptr, d_hist ? d_hist + (size_t)i0 * 768 : nullptr, st)) return rc;
      }
    }
    return 0;
  };
  if (int rc = run()) {
    if (host) drain_streams(h, st);             // no copy into / out of the caller's buffers is left in flight
    return rc;
  }
  if (host) { CK(h, cudaStreamSynchronize(h->d2h_stream)); CK(h, cudaStreamSynchronize(st)); if (int rc = check_device_error(h)) return rc; }
  return NNIC_OK;
}

int nnic_encode(nnic_t* h, const uint8_t* rgb, int N, int H, int W, uint8_t* latent, float* prequant, int mem_kind,
                void* stream) {
  return encode_impl(h, rgb, N, H, W, latent, prequant, nullptr, mem_kind, stream);
}

int nnic_decode(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, uint8_t* rgb, float* prequant, int mem_kind,
                void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!rgb || !latent) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_decode: NULL buffer");
  if (N <= 0 || lh <= 0 || lw <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_decode: non-positive shape %dx%dx%d", N, lh, lw);
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"latent", latent}, {"rgb", rgb}, {"prequant", prequant}})) return rc;
  if (int rc = check_device_error(h)) return rc;
  if (int rc = finalize_weights(h, 1)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const size_t img_px = (size_t)lh * lw * 64, lat_px = (size_t)lh * lw;
  const bool host = mem_kind == NNIC_MEM_HOST;
  int mb = pick_micro_batch(h, N, img_px);
  if (host && h->micro_batch == 0 && N >= 2) {
    const int quarter = (N + 3) / 4;
    if (quarter < mb) mb = quarter;
  }
  size_t need = dec_act_need(split, 3 * (size_t)mb, lh, lw);
  if (host) need += 2 * (pad1k(mb * img_px * 3) + pad1k(mb * lat_px * 96) + (prequant ? pad1k(mb * img_px * 3 * 4) : 0)) + 8192;
  if (int rc = ensure_buf(h, h->arena, need)) return rc;
  h->arena_used = 0;
  uint8_t *d_rgb[2] = {nullptr, nullptr}, *d_lat[2] = {nullptr, nullptr}; float* d_pre[2] = {nullptr, nullptr};
  if (host) {
    for (int s2 = 0; s2 < 2; ++s2) {
      d_rgb[s2] = (uint8_t*)arena_take(h, mb * img_px * 3);
      d_lat[s2] = (uint8_t*)arena_take(h, mb * lat_px * 96);
      if (prequant) d_pre[s2] = (float*)arena_take(h, mb * img_px * 3 * 4);
    }
  }
  auto run = [&]() -> int {
    int idx = 0;
    for (int i0 = 0; i0 < N; i0 += mb, ++idx) {
      const int nb = (N - i0) < mb ? (N - i0) : mb;
      const uint8_t* src = latent + (size_t)i0 * lat_px * 96;
      uint8_t* dst = rgb + (size_t)i0 * img_px * 3;
      float* pre = prequant ? prequant + (size_t)i0 * img_px * 3 : nullptr;
      if (host) {
        const int s2 = idx & 1;
        if (idx >= 2) CK(h, cudaStreamWaitEvent(h->h2d_stream, h->ev_out[s2], 0));
        CK(h, cudaMemcpyAsync(d_lat[s2], src, nb * lat_px * 96, cudaMemcpyHostToDevice, h->h2d_stream));
        CK(h, cudaEventRecord(h->ev_in[s2], h->h2d_stream));
        CK(h, cudaStreamWaitEvent(st, h->ev_in[s2], 0));
        if (int rc = decode_batch(h, d_lat[s2], nullptr, nb, lh, lw, d_rgb[s2], d_pre[s2], nullptr, st)) return rc;
        CK(h, cudaEventRecord(h->ev_comp[s2], st));
        CK(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_comp[s2], 0));
        CK(h, cudaMemcpyAsync(dst, d_rgb[s2], nb * img_px * 3, cudaMemcpyDeviceToHost, h->d2h_stream));
        if (pre) CK(h, cudaMemcpyAsync(pre, d_pre[s2], nb * img_px * 3 * 4, cudaMemcpyDeviceToHost, h->d2h_stream));
        CK(h, cudaEventRecord(h->ev_out[s2], h->d2h_stream));
      } else {
        if (int rc = decode_batch(h, src, nullptr, nb, lh, lw, dst, pre, nullptr, st)) return rc;
      }
    }
    return 0;
  };
  if (int rc = run()) {
    if (host) drain_streams(h, st);             // no copy into / out of the caller's buffers is left in flight
    return rc;
  }
  if (host) { CK(h, cudaStreamSynchronize(h->d2h_stream)); CK(h, cudaStreamSynchronize(st)); if (int rc = check_device_error(h)) return rc; }
  return NNIC_OK;
}

int nnic_run_encoder_planes(nnic_t* h, const float* planes, int N, int H, int W, float* out, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!planes || !out) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_run_encoder_planes: NULL buffer");
  if (N <= 0 || H <= 0 || W <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "non-positive shape");
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"planes", planes}, {"out", out}})) return rc;
  if (int rc = check_device_error(h)) return rc;
  if (int rc = finalize_weights(h, 0)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const int lh = (H + 7) / 8, lw = (W + 7) / 8;
  const size_t in_elems = (size_t)3 * N * H * W, out_elems = (size_t)3 * N * lh * lw * 32;
  const bool host = mem_kind == NNIC_MEM_HOST;
  size_t need = enc_act_need(split, 3 * (size_t)N, H, W) + (host ? pad1k(in_elems * 4) + pad1k(out_elems * 4) + 4096 : 0);
  if (int rc = ensure_buf(h, h->arena, need)) return rc;
  h->arena_used = 0;
  const float* d_in = planes; float* d_out = out;
  if (host) {
    float* t_in = (float*)arena_take(h, in_elems * 4);
    d_out = (float*)arena_take(h, out_elems * 4);
    CK(h, cudaMemcpyAsync(t_in, planes, in_elems * 4, cudaMemcpyHostToDevice, st));
    d_in = t_in;
  }
  if (int rc = encode_batch(h, nullptr, d_in, N, H, W, nullptr, nullptr, d_out, nullptr, st)) return rc;
  if (host) {
    CK(h, cudaMemcpyAsync(out, d_out, out_elems * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    if (int rc = check_device_error(h)) return rc;
  }
  return NNIC_OK;
}

int nnic_run_decoder_planes(nnic_t* h, const float* planes, int N, int lh, int lw, float* out, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!planes || !out) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_run_decoder_planes: NULL buffer");
  if (N <= 0 || lh <= 0 || lw <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "non-positive shape");
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"planes", planes}, {"out", out}})) return rc;
  if (int rc = check_device_error(h)) return rc;
  if (int rc = finalize_weights(h, 1)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const size_t in_elems = (size_t)3 * N * lh * lw * 32, out_elems = (size_t)3 * N * lh * lw * 64;
  const bool host = mem_kind == NNIC_MEM_HOST;
  size_t need = dec_act_need(split, 3 * (size_t)N, lh, lw) + (host ? pad1k(in_elems * 4) + pad1k(out_elems * 4) + 4096 : 0);
  if (int rc = ensure_buf(h, h->arena, need)) return rc;
  h->arena_used = 0;
  const float* d_in = planes; float* d_out = out;
  if (host) {
    float* t_in = (float*)arena_take(h, in_elems * 4);
    d_out = (float*)arena_take(h, out_elems * 4);
    CK(h, cudaMemcpyAsync(t_in, planes, in_elems * 4, cudaMemcpyHostToDevice, st));
    d_in = t_in;
  }
  if (int rc = decode_batch(h, nullptr, d_in, N, lh, lw, nullptr, nullptr, d_out, st)) return rc;
  if (host) {
    CK(h, cudaMemcpyAsync(out, d_out, out_elems * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    if (int rc = check_device_error(h)) return rc;
  }
  return NNIC_OK;
}

// Rate outputs: device scratch (host callers) or the caller's device buffers.
struct RateBufs {
  uint32_t* d_hist; float* d_ent; float* d_bpp; unsigned long long* d_glob; uint8_t* d_lat;
};
static int rate_setup(nnic_t* h, int N, size_t lat_bytes_host, uint32_t* hist, float* entropy_bits, float* bpp,
                      uint64_t* hist_global, bool host, cudaStream_t st, RateBufs& rb) {
  const size_t hist_bytes = (size_t)N * 768 * 4;
  size_t need = pad1k(hist_bytes) + pad1k((size_t)N * 3 * 4) + pad1k((size_t)N * 4) + pad1k(768 * 8) + pad1k(lat_bytes_host) + 8192;
  if (int rc = ensure_buf(h, h->rate_scratch, need)) return rc;
  uint8_t* base = (uint8_t*)h->rate_scratch.ptr;
  size_t off = 0;
  auto take = [&](size_t b) { void* p = base + off; off += pad1k(b); return p; };
  rb.d_hist = (uint32_t*)take(hist_bytes);
  rb.d_ent = (float*)take((size_t)N * 3 * 4);
  rb.d_bpp = (float*)take((size_t)N * 4);
  rb.d_glob = (unsigned long long*)take(768 * 8);
  rb.d_lat = lat_bytes_host ? (uint8_t*)take(lat_bytes_host) : nullptr;
  if (!host) {
    if (hist) rb.d_hist = hist;
    if (entropy_bits) rb.d_ent = entropy_bits;
    if (bpp) rb.d_bpp = bpp;
    if (hist_global) rb.d_glob = (unsigned long long*)hist_global;
  }
  CK(h, cudaMemsetAsync(rb.d_hist, 0, hist_bytes, st));
  return 0;
}
// entropy / bpp / global counts from rb.d_hist, and the copies back for host callers
static int rate_finish(nnic_t* h, int N, int lh, int lw, int H, int W, uint32_t* hist, float* entropy_bits, float* bpp,
                       uint64_t* hist_global, bool host, cudaStream_t st, const RateBufs& rb) {
  const size_t hist_bytes = (size_t)N * 768 * 4;
  if (entropy_bits || bpp)
    CKL(h, K_ENTROPY, st, launch_entropy_u32(rb.d_hist, N, (float)((size_t)lh * lw * 32), (float)((size_t)H * W), rb.d_ent, rb.d_bpp, st));
  if (hist_global) {
    if (host) CK(h, cudaMemcpyAsync(rb.d_glob, hist_global, 768 * 8, cudaMemcpyHostToDevice, st));
    CKL(h, K_HIST_REDUCE, st, launch_hist_reduce(rb.d_hist, N, rb.d_glob, st));
  }
  if (host) {
    if (hist) CK(h, cudaMemcpyAsync(hist, rb.d_hist, hist_bytes, cudaMemcpyDeviceToHost, st));
    if (entropy_bits) CK(h, cudaMemcpyAsync(entropy_bits, rb.d_ent, (size_t)N * 3 * 4, cudaMemcpyDeviceToHost, st));
    if (bpp) CK(h, cudaMemcpyAsync(bpp, rb.d_bpp, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    if (hist_global) CK(h, cudaMemcpyAsync(hist_global, rb.d_glob, 768 * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
  }
  return NNIC_OK;
}

int nnic_rate(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, int H, int W, uint32_t* hist, float* entropy_bits,
              float* bpp, uint64_t* hist_global, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!latent) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_rate: latent is NULL");
  if (N <= 0 || lh <= 0 || lw <= 0 || H <= 0 || W <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_rate: non-positive shape");
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"latent", latent}, {"hist", hist}, {"entropy_bits", entropy_bits}, {"bpp", bpp},
                                               {"hist_global", hist_global}})) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool host = mem_kind == NNIC_MEM_HOST;
  const size_t lat_bytes = (size_t)N * lh * lw * 96;
  RateBufs rb;
  if (int rc = rate_setup(h, N, host ? lat_bytes : 0, hist, entropy_bits, bpp, hist_global, host, st, rb)) return rc;
  const uint8_t* d_lat = latent;
  if (host) {
    CK(h, cudaMemcpyAsync(rb.d_lat, latent, lat_bytes, cudaMemcpyHostToDevice, st));
    d_lat = rb.d_lat;
  }
  CKL(h, K_HIST, st, launch_hist(d_lat, N, (size_t)lh * lw, rb.d_hist, h->num_sms, h->hist_variant, st));
  return rate_finish(h, N, lh, lw, H, W, hist, entropy_bits, bpp, hist_global, host, st, rb);
}

int nnic_encode_rate(nnic_t* h, const uint8_t* rgb, int N, int H, int W, uint8_t* latent, uint32_t* hist,
                     float* entropy_bits, float* bpp, uint64_t* hist_global, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (N <= 0 || H <= 0 || W <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_encode_rate: non-positive shape %dx%dx%d", N, H, W);
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"hist", hist}, {"entropy_bits", entropy_bits}, {"bpp", bpp}, {"hist_global", hist_global}}))
    return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool host = mem_kind == NNIC_MEM_HOST;
  RateBufs rb;
  if (int rc = rate_setup(h, N, 0, hist, entropy_bits, bpp, hist_global, host, st, rb)) return rc;
  if (int rc = encode_impl(h, rgb, N, H, W, latent, nullptr, rb.d_hist, mem_kind, stream)) return rc;
  return rate_finish(h, N, (H + 7) / 8, (W + 7) / 8, H, W, hist, entropy_bits, bpp, hist_global, host, st, rb);
}

int nnic_rate_channels(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, uint64_t* hist_channels, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!latent || !hist_channels) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_rate_channels: NULL buffer");
  if (N <= 0 || lh <= 0 || lw <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_rate_channels: non-positive shape");
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"latent", latent}, {"hist_channels", hist_channels}})) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pixels = (size_t)N * lh * lw, tab = (size_t)96 * 256 * 8;
  if (mem_kind == NNIC_MEM_DEVICE) {
    CKL(h, K_HIST, st, launch_hist_channels(latent, pixels, (unsigned long long*)hist_channels, h->num_sms, st));
    return NNIC_OK;
  }
  if (int rc = ensure_buf(h, h->rate_scratch, pad1k(tab) + pad1k(pixels * 96) + 4096)) return rc;
  unsigned long long* d_tab = (unsigned long long*)h->rate_scratch.ptr;
  uint8_t* d_lat = (uint8_t*)h->rate_scratch.ptr + pad1k(tab);
  CK(h, cudaMemcpyAsync(d_tab, hist_channels, tab, cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpyAsync(d_lat, latent, pixels * 96, cudaMemcpyHostToDevice, st));
  CKL(h, K_HIST, st, launch_hist_channels(d_lat, pixels, d_tab, h->num_sms, st));
  CK(h, cudaMemcpyAsync(hist_channels, d_tab, tab, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return NNIC_OK;
}

int nnic_hist_allreduce(nnic_t* h, void* nccl_comm, uint64_t* hist_global, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!nccl_comm || !hist_global) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_hist_allreduce: NULL communicator or buffer");
  // int ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static allreduce_fn fn = nullptr;
  if (!fn) {
    fn = (allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!fn) {
      void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (lib) fn = (allreduce_fn)dlsym(lib, "ncclAllReduce");
    }
    if (!fn) return fail(h, NNIC_ERR_CUDA, "nnic_hist_allreduce: NCCL (libnccl.so.2) is not available in this process");
  }
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, NNIC_MEM_DEVICE, {{"hist_global", hist_global}})) return rc;
  constexpr int kNcclUint64 = 5, kNcclSum = 0;       // nccl.h: ncclDataType_t / ncclRedOp_t
  const int rc = fn(hist_global, hist_global, 768, kNcclUint64, kNcclSum, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return fail(h, NNIC_ERR_CUDA, "ncclAllReduce failed with ncclResult_t %d", rc);
  return NNIC_OK;
}

int nnic_entropy_from_counts(nnic_t* h, const uint64_t* counts, int rows, float* entropy_bits, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!counts || !entropy_bits || rows <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_entropy_from_counts: bad argument");
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"counts", counts}, {"entropy_bits", entropy_bits}})) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mem_kind == NNIC_MEM_DEVICE) {
    CKL(h, K_ENTROPY, st, launch_entropy_u64((const unsigned long long*)counts, rows, entropy_bits, st));
    return NNIC_OK;
  }
  size_t need = pad1k((size_t)rows * 256 * 8) + pad1k((size_t)rows * 4) + 4096;
  if (int rc = ensure_buf(h, h->rate_scratch, need)) return rc;
  unsigned long long* d_c = (unsigned long long*)h->rate_scratch.ptr;
  float* d_e = (float*)((uint8_t*)h->rate_scratch.ptr + pad1k((size_t)rows * 256 * 8));
  CK(h, cudaMemcpyAsync(d_c, counts, (size_t)rows * 256 * 8, cudaMemcpyHostToDevice, st));
  CKL(h, K_ENTROPY, st, launch_entropy_u64(d_c, rows, d_e, st));
  CK(h, cudaMemcpyAsync(entropy_bits, d_e, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return NNIC_OK;
}

// ---- forward-only extras of the reference's training step -----------------------------------------

int nnic_entropynet_set_weights(nnic_t* h, int layer, const float* kernel, const float* bias, int features) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (layer < 0 || layer > 4 || !kernel || !bias) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_entropynet_set_weights: bad layer %d or NULL buffer", layer);
  size_t nk, nb;
  if (layer < 3) { nk = (size_t)kEnt[layer].k * kEnt[layer].k * kEnt[layer].cin * kEnt[layer].cout; nb = kEnt[layer].cout; }
  else if (layer == 3) {
    if (features <= 0 || features % 64) return fail(h, NNIC_ERR_INVALID_ARG, "dense1 needs features = 64 * ceil(h/2) * ceil(w/2), got %d", features);
    nk = (size_t)features * 512; nb = 512; h->ent_features = features;
  } else { nk = 512; nb = 1; }
  h->ent_kernel[layer].assign(kernel, kernel + nk);
  h->ent_bias[layer].assign(bias, bias + nb);
  h->ent_have[layer] = true;
  h->ent_dirty = true;
  return NNIC_OK;
}

int nnic_entropynet_forward(nnic_t* h, const float* encoded, int P, int lh, int lw, float* approx_entropy, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!encoded || !approx_entropy) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_entropynet_forward: NULL buffer");
  if (P <= 0 || lh <= 0 || lw <= 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_entropynet_forward: non-positive shape");
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"encoded", encoded}, {"approx_entropy", approx_entropy}})) return rc;
  if (int rc = check_device_error(h)) return rc;
  if (int rc = finalize_entropynet(h)) return rc;
  const int h2 = (lh + 1) / 2, w2 = (lw + 1) / 2;
  const int F = h2 * w2 * 64;
  if (F != h->ent_features)
    return fail(h, NNIC_ERR_SHAPE, "Entropynet.dense1 was built for %d features; a %dx%d latent flattens to %d (training.py:30-31: Flatten -> Dense(512))", h->ent_features, lh, lw, F);
  cudaStream_t st = (cudaStream_t)stream;
  // the stride-2 first layer reads its input through the parity view (even extents): odd latent sizes take the FFMA kernels
  const int saved_arith = h->arith;
  if ((lh | lw) & 1) h->arith = NNIC_ARITH_SIMT_F32;
  const bool split = h->arith == NNIC_ARITH_TC_SPLIT;
  const bool host = mem_kind == NNIC_MEM_HOST;
  const size_t in_elems = (size_t)P * lh * lw * 32;
  size_t need = act_bytes(split, P, lh, lw, 32) + 3 * act_bytes(split, P, h2, w2, 64) + pad1k((size_t)P * 512 * 4) + pad1k((size_t)P * 4) + 16384;
  if (host) need += pad1k(in_elems * 4);
  int rc = ensure_buf(h, h->arena, need);
  if (rc) { h->arith = saved_arith; return rc; }
  h->arena_used = 0;
  auto run = [&]() -> int {
    const float* d_in = encoded;
    if (host) {
      float* t = (float*)arena_take(h, in_elems * 4);
      CK(h, cudaMemcpyAsync(t, encoded, in_elems * 4, cudaMemcpyHostToDevice, st));
      d_in = t;
    }
    Act a0 = take_act(h, split, P, lh, lw, 32);
    Act a1 = take_act(h, split, P, h2, w2, 64), a2 = take_act(h, split, P, h2, w2, 64), a3 = take_act(h, split, P, h2, w2, 64);
    float* hid = (float*)arena_take(h, (size_t)P * 512 * 4);
    float* d_out = host ? (float*)arena_take(h, (size_t)P * 4) : approx_entropy;
    if (split) CKL(h, K_F32_SPLIT, st, launch_f32_to_split(d_in, in_elems, a0.hi, a0.lo, st));
    else a0.f32 = const_cast<float*>(d_in);
    if (int r2 = run_gemm_layer(h, 2, 0, a0, a1, nullptr, P, P, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return r2;
    if (int r2 = run_gemm_layer(h, 2, 1, a1, a2, nullptr, P, P, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return r2;
    if (int r2 = run_gemm_layer(h, 2, 2, a2, a3, nullptr, P, P, TC_OUT_SPLIT, nullptr, nullptr, nullptr, st)) return r2;
    CKL(h, K_DENSE, st, launch_dense512(a3.hi, a3.lo, a3.f32, P, F, h->ent_d1_w, h->ent_d1_b, hid, st));
    CKL(h, K_DENSE, st, launch_dense1_clip(hid, P, h->ent_d2_w, h->ent_d2_b, d_out, st));
    if (host) {
      CK(h, cudaMemcpyAsync(approx_entropy, d_out, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
      CK(h, cudaStreamSynchronize(st));
      if (int r2 = check_device_error(h)) return r2;
    }
    return 0;
  };
  rc = run();
  h->arith = saved_arith;
  if (rc && host) cudaStreamSynchronize(st);
  return rc;
}

int nnic_noise_quantise(nnic_t* h, const float* encoded, size_t count, uint64_t seed, const float* noise, float* out, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!encoded || !out || count == 0) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_noise_quantise: NULL buffer or empty input");
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"encoded", encoded}, {"noise", noise}, {"out", out}})) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mem_kind == NNIC_MEM_DEVICE) {
    CKL(h, K_NOISE, st, launch_noise_quantise(encoded, count, seed, noise, out, h->num_sms, st));
    return NNIC_OK;
  }
  if (int rc = ensure_buf(h, h->arena, (noise ? 3 : 2) * pad1k(count * 4) + 8192)) return rc;
  h->arena_used = 0;
  float* d_x = (float*)arena_take(h, count * 4);
  float* d_o = (float*)arena_take(h, count * 4);
  float* d_n = noise ? (float*)arena_take(h, count * 4) : nullptr;
  CK(h, cudaMemcpyAsync(d_x, encoded, count * 4, cudaMemcpyHostToDevice, st));
  if (noise) CK(h, cudaMemcpyAsync(d_n, noise, count * 4, cudaMemcpyHostToDevice, st));
  CKL(h, K_NOISE, st, launch_noise_quantise(d_x, count, seed, d_n, d_o, h->num_sms, st));
  CK(h, cudaMemcpyAsync(out, d_o, count * 4, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return NNIC_OK;
}

int nnic_ssim(nnic_t* h, const float* a, const float* b, int P, int H, int W, float* ssim, int mem_kind, void* stream) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  if (!a || !b || !ssim) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_ssim: NULL buffer");
  if (P <= 0 || H < 11 || W < 11) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_ssim: images must be at least 11 x 11 (the window of tf.image.ssim), got %d x %d x %d", P, H, W);
  if (mem_kind != NNIC_MEM_HOST && mem_kind != NNIC_MEM_DEVICE) return fail(h, NNIC_ERR_INVALID_ARG, "bad mem_kind %d", mem_kind);
  DeviceGuard g(h->device);
  if (int rc = check_device_ptrs(h, mem_kind, {{"a", a}, {"b", b}, {"ssim", ssim}})) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool host = mem_kind == NNIC_MEM_HOST;
  const size_t elems = (size_t)P * H * W, npart = ssim_partial_count(P, H, W);
  if (int rc = ensure_buf(h, h->rate_scratch, pad1k(npart * 4) + (host ? 2 * pad1k(elems * 4) + pad1k((size_t)P * 4) : 0) + 8192)) return rc;
  uint8_t* base = (uint8_t*)h->rate_scratch.ptr;
  float* d_part = (float*)base;
  const float *d_a = a, *d_b = b;
  float* d_out = ssim;
  if (host) {
    float* ta = (float*)(base + pad1k(npart * 4));
    float* tb = (float*)(base + pad1k(npart * 4) + pad1k(elems * 4));
    d_out = (float*)(base + pad1k(npart * 4) + 2 * pad1k(elems * 4));
    CK(h, cudaMemcpyAsync(ta, a, elems * 4, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(tb, b, elems * 4, cudaMemcpyHostToDevice, st));
    d_a = ta; d_b = tb;
  }
  CKL(h, K_SSIM, st, launch_ssim(d_a, d_b, P, H, W, d_part, d_out, st));
  if (host) {
    CK(h, cudaMemcpyAsync(ssim, d_out, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
  }
  return NNIC_OK;
}

int nnic_set_profiling(nnic_t* h, int on) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  h->prof = on != 0;
  return NNIC_OK;
}

// Sum of the event-timed durations (ms) and launch counts per kernel id since the last collect.
int nnic_profile_collect(nnic_t* h, float* ms_per_kernel, int* launches_per_kernel, int capacity) {
  if (!h || !ms_per_kernel || !launches_per_kernel || capacity < K_COUNT) return fail(h, NNIC_ERR_INVALID_ARG, "nnic_profile_collect: need capacity >= %d", (int)K_COUNT);
  DeviceGuard g(h->device);
  for (int i = 0; i < capacity; ++i) { ms_per_kernel[i] = 0.f; launches_per_kernel[i] = 0; }
  for (auto& r : h->prof_recs) {
    CK(h, cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CK(h, cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_per_kernel[r.id] += ms;
    launches_per_kernel[r.id] += 1;
    h->prof_pool.push_back(r.e0);
    h->prof_pool.push_back(r.e1);
  }
  h->prof_recs.clear();
  return K_COUNT;
}

// Debug aid for the parity tests: copy an intermediate activation of the most recent encode
// (slots 0-3: conv1, conv2, conv3, conv4+res) or decode (slots 4-7: latent/255, dconv1, dconv5,
// dconv6+res) micro-batch to host as fp32.  Returns the element count, or a negative status.
long long nnic_debug_fetch(nnic_t* h, int slot, float* out, long long capacity) {
  if (!h || slot < 0 || slot > 7) return NNIC_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  const nnic_handle::Dbg& d = h->dbg[slot];
  if (!d.count) return 0;
  if (!out) return (long long)d.count;
  if ((long long)d.count > capacity) return fail(h, NNIC_ERR_INVALID_ARG, "debug buffer too small");
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "sync failed");
  if (d.f32) {
    if (cudaMemcpy(out, d.f32, d.count * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "copy failed");
  } else {
    std::vector<__half> hi(d.count), lo(d.count);
    if (cudaMemcpy(hi.data(), d.hi, d.count * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "copy failed");
    if (cudaMemcpy(lo.data(), d.lo, d.count * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "copy failed");
    for (size_t i = 0; i < d.count; ++i) out[i] = (__half2float(hi[i]) + __half2float(lo[i])) * ACT_INV_SCALE;
  }
  return (long long)d.count;
}

// The split-fp16 activation planes hold v * 16 and saturate at the fp16 maximum (|v| > 4094), where the fp32 reference would
// carry on: a trained codec stays far below that (activations of O(1)), but nothing in the kernels reports it.  This debug
// call counts the saturated values in the activations of the most recent encode (conv1 .. conv4) and decode (dconv1 .. dconv6, and
// dconv7 when its output is stored: NNIC_FUSE_D78=0)
// micro-batch of the handle; 0 means the split representation was exact to its 22 bits everywhere.
long long nnic_debug_saturated(nnic_t* h) {
  if (!h) return NNIC_ERR_INVALID_ARG;
  DeviceGuard g(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "sync failed");
  unsigned long long* d_cnt = nullptr;
  if (cudaMalloc(&d_cnt, 8) != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "cudaMalloc failed");
  cudaMemset(d_cnt, 0, 8);
  for (int s = 0; s < 9; ++s) {
    const nnic_handle::Dbg& d = h->dbg[s];
    if (!d.hi || !d.stored || (s == 4 && h->int_latent && !h->decode_fp16)) continue;   // slot 4 may hold the integer symbols
    if (launch_count_saturated(d.hi, d.stored, d_cnt, h->num_sms, nullptr) != cudaSuccess) { cudaFree(d_cnt); return fail(h, NNIC_ERR_CUDA, "launch failed"); }
  }
  unsigned long long cnt = 0;
  const cudaError_t e = cudaMemcpy(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost);
  cudaFree(d_cnt);
  if (e != cudaSuccess) return fail(h, NNIC_ERR_CUDA, "copy failed");
  return (long long)cnt;
}

}  // extern "C"
