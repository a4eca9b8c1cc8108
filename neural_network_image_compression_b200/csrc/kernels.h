// Host-side launch interface of the libnnic kernels.  Everything here enqueues on `stream` and
// returns the cudaError_t of the launch.  Layout conventions:
//   plane batches are plane-major: P = 3N, p = plane*N + n; planes p < n_split use weight set 0
//   (the 'Y' network), the others weight set 1 ('CbCr')  -- reference tf2_0/src/utils.py:19-24.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace nnic {

// Kernel attributes (opt-in shared memory size) are per device: true the first time it is called on the current device.
inline bool first_use_on_device(unsigned long long& seen_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen_mask & bit) return false;
  seen_mask |= bit;
  return true;
}


// Programmatic dependent launch (NNIC_PDL=0 switches it off): a kernel launched through launch_kernel(..., pdl = true) may start
// while the previous kernel of the stream is still running -- its prologue (barrier init, TMEM allocation, weight and bias
// staging) overlaps the predecessor's tail -- and must execute pdl_wait() (common.cuh) before it touches anything an earlier
// kernel writes or reads.  Every such kernel also executes pdl_trigger() at its start.
extern int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

struct ColourConsts {
  float k[3][3];     // RGB -> YCbCr rows   (float)(ycbcr_kernel)      utils.py:7
  float kinv[3][3];  // YCbCr -> RGB rows   (float)(inv(ycbcr_kernel)) utils.py:8
  float off[3];      // (float)(ycbcr_off)                             utils.py:9
};
const ColourConsts& colour_consts();

// ---- conv1: [colour +] Conv2D(1->32, 5x5, s2, SAME) + bias + leaky ----------------------------
// in_rgb: u8 [N,H,W,3] (colour transform fused) or in_planes: f32 [3N,H,W,1].  Output [3N,Ho,Wo,32]
// as split fp16 (out_hi/out_lo) or fp32 (out_f32).  w: HOST [2][25][32] tap-major, bias HOST [2][32]
// (they travel to the kernel as __grid_constant__ parameters).
cudaError_t launch_conv1(const uint8_t* in_rgb, const float* in_planes, int N, int H, int W,
                         const float* w, const float* bias, __half* out_hi, __half* out_lo,
                         float* out_f32, cudaStream_t stream);

// ---- generic fp32 FFMA convolution over a tap program (cross-check path) ----------------------
// in f32 [P,Hi,Wi,CIN], out f32 [P,Ho,Wo,COUT]; w [2][ntaps_total][CIN][COUT]; bias [2][COUT];
// res optional f32 [P,Ho,Wo,COUT] added after the activation; Hp x Wp = phase grid.
cudaError_t launch_simt_conv(int cin, int cout, const float* in, int P, int Hi, int Wi, float* out,
                             int Ho, int Wo, int Hp, int Wp, const float* w, int ntaps_total,
                             const float* bias, const float* res, const SimtJobs& jobs, int n_split,
                             int clamp01, cudaStream_t stream);

// ---- dconv8: Conv2DTranspose(64->1, 5x5, s2, SAME) + bias + leaky + clip, then
//      convert_to_rgb + clip + *255 + round + uint8 pack, all three planes of an image per block ---
// input [3N,Hi,Wi,64] split fp16 or f32; w HOST [2][25][64] tap-major; bias HOST [2][1].
// Outputs (each optional): rgb u8 [N,2Hi,2Wi,3]; prequant f32 [N,2Hi,2Wi,3]; planes f32 [3][N,2Hi,2Wi,1].
cudaError_t launch_dconv8(const __half* in_hi, const __half* in_lo, const float* in_f32, int N, int Hi,
                          int Wi, const float* w, const float* bias, uint8_t* rgb, float* prequant,
                          float* planes, cudaStream_t stream);

// ---- latent u8 [N,lh,lw,96] -> /255 -> planes [3N,lh,lw,32] (split fp16 or f32) -----------------
// integer_symbols: out_hi receives the symbols themselves (0..255, exact in fp16) and out_lo is not written
cudaError_t launch_latent_expand(const uint8_t* latent, int N, int lh, int lw, __half* out_hi,
                                 __half* out_lo, float* out_f32, bool integer_symbols, cudaStream_t stream);
// f32 -> split fp16, elementwise (count elements, multiple of 4)
cudaError_t launch_f32_to_split(const float* in, size_t count, __half* out_hi, __half* out_lo,
                                cudaStream_t stream);
// clipped f32 planes [3N,lh,lw,32] -> u8 latent [N,lh,lw,96] (+ optional f32 prequant [N,lh,lw,96])
cudaError_t launch_quantise(const float* planes, int N, int lh, int lw, uint8_t* latent, float* prequant,
                            cudaStream_t stream);

// adds to *out the number of values of a split-fp16 hi plane at the saturation limit (|v * 16| >= 65504)
cudaError_t launch_count_saturated(const __half* hi, size_t n, unsigned long long* out, int num_sms, cudaStream_t stream);

// ---- rate ---------------------------------------------------------------------------------------
// hist u32 [N][3][256] must be zeroed by the caller (launch_hist adds into it).
cudaError_t launch_hist(const uint8_t* latent, int N, size_t pixels_per_image, uint32_t* hist, int num_sms, int variant,
                        cudaStream_t stream);
// hist_ch u64 [96][256], added to: counts per latent feature channel over `total_pixels` latent pixels (all images)
cudaError_t launch_hist_channels(const uint8_t* latent, size_t total_pixels, unsigned long long* hist_ch, int num_sms,
                                 cudaStream_t stream);
cudaError_t launch_hist_reduce(const uint32_t* hist, int N, unsigned long long* hist_global,
                               cudaStream_t stream);
// entropy[N][3] from hist u32; bpp[N] = sum_p entropy * symbols_per_plane / pixels  (optional)
cudaError_t launch_entropy_u32(const uint32_t* hist, int N, float symbols_per_plane, float pixels,
                               float* entropy, float* bpp, cudaStream_t stream);
cudaError_t launch_entropy_u64(const unsigned long long* counts, int rows, float* entropy,
                               cudaStream_t stream);

// ---- forward-only extras of the reference's training step (train_extras.cu; tf2_0/src/training.py:25-42, 87-88, 108-119) ----
// Dense(512) on [P][F] inputs given as split fp16 planes (x_hi / x_lo) or fp32 (x_f32); Wt is the Keras kernel [F][512]
cudaError_t launch_dense512(const __half* x_hi, const __half* x_lo, const float* x_f32, int P, int F, const float* Wt,
                            const float* bias, float* out, cudaStream_t stream);
// Dense(1) + clip(0, 8): hid [P][512] -> out [P]
cudaError_t launch_dense1_clip(const float* hid, int P, const float* w, float b, float* out, cudaStream_t stream);
// out = clip(x + u/255, 0, 1); u from `noise` (optional) or Philox4x32-10(seed, element index)
cudaError_t launch_noise_quantise(const float* x, size_t count, unsigned long long seed, const float* noise, float* out,
                                  int num_sms, cudaStream_t stream);
// tf.image.ssim(a, b, max_val=1) of fp32 [P][H][W] images; partial: scratch of ssim_partial_count floats
cudaError_t launch_ssim(const float* a, const float* b, int P, int H, int W, float* partial, float* out, cudaStream_t stream);
size_t ssim_partial_count(int P, int H, int W);

// ---- tensor-core convolutions ------------------------------------------------------------------------
enum TcOutMode { TC_OUT_SPLIT = 0, TC_OUT_F32 = 1, TC_OUT_QUANT = 2 };

// ---- tensor-core convolution on a halo patch (tc_conv_patch.cu): every GEMM-shaped layer; all taps lie within the
//      3x3 neighbourhood of the tile in the view the layer reads ---------------------------------------------------
struct TcPatchStep {
  uint32_t a_off;      // byte offset of the tap's first pixel row inside the hi patch (tc_patch_a_offset)
  int16_t w_row;       // first row of this tap's [COUT x K] tile in the weight matrix
  int16_t pad_;
};
struct TcPatchJob {
  int nsteps;
  int nchains;         // accumulation chains (TMEM slots) per tile: three taps each, never across patches
  int out_oy, out_ox;
  TcPatchStep steps[MAX_STEPS];
};
struct TcPatchParams {
  int njobs;
  TcPatchJob jobs[MAX_JOBS];
  // activation patches per work item: 1, or 2 for conv2 (one per input-row parity; job 0's first seg_steps[0] steps
  // read patch 0, the rest patch 1).  a_off bit 31 = the step only uses the upper half of its K slab.
  // conv8 uses four (row parity, column parity) patches; patch_c0 = first element of the view's inner dimension.
  int npatch;
  int patch_py[4];
  int patch_c0[4];
  int seg_steps[4];
  int cout;                   // 64, or 32 (conv8)
  int fast;                   // one fp16 product per MAC, hi planes only (decoder, nnic_set_decode_precision)
  int cluster;                // 2: CTA pairs share every weight tile through TMA multicast; else 1
  int pin;                    // nine-tap layers: seven weight tiles stay in shared memory for all items of a weight set (tc_conv_patch.cu PIN)
  int a_hi_only;              // the input is exact in its hi plane (integer latent symbols): no lo plane, no A_lo x W_hi product
  unsigned long long wait_timeout;   // bound of a barrier wait in SM cycles, 0 = unbounded (tc_common.cuh WaitCtx)
  int kernel_tag;             // reported with a timed-out wait: kernel id of the launch + 1 (NNIC_KERNEL_*)
  int P, n_split;
  int Hp, Wp;
  int Ho, Wo, out_stride;
  int Hs, Ws;                 // storage rows / columns per plane of the output and the residual (>= Ho, Wo)
  int rows_per_set;
  float inv_scale[2];
  const float* bias;
  const __half* res_hi;
  const __half* res_lo;
  int out_mode;               // TC_OUT_SPLIT; cout == 32 also TC_OUT_F32 / TC_OUT_QUANT
  int clamp01;                // clip the activation to [0,1] (conv8: encoder.py:32)
  uint8_t* out_u8;            // TC_OUT_QUANT: latent [N,Ho,Wo,96], N = P/3
  float* out_prequant;        // TC_OUT_QUANT, optional: f32 [N,Ho,Wo,96]
  uint32_t* hist;             // TC_OUT_QUANT, optional: [N][3][256] symbol counts, added to (tf1_13/src/training.py:62-68)
  long long* dbg_buf;         // development: per-CTA role timers [grid][4][8] (NNIC_TC_PROF)
  int dbg;                    // development switches (NNIC_TC_DBG): 1 skip MMAs, 2 skip stores, 4 skip W loads, 8 skip TMEM loads, 16 skip patch loads, 32 / 64 load W / patches only once
  __half* out_hi;
  __half* out_lo;
  float* out_f32;
  // dconv7 fused with dconv8's response GEMM (decoder.py:16-17): instead of the activation, the layer writes
  // f8_out = R [P][tile][25 taps][4 output phases][128 tile pixels] fp32, R = x . K8[tap] per dconv7 output pixel, tiles of 16 x 8
  // pixels of the Hp x Wp input grid in row-major order (read by launch_dconv8_gather);
  // f8_w_hi / f8_w_lo: dconv8's [2 sets][32 taps (25 used)][64] fp16 matrices
  float* f8_out;
  const __half* f8_w_hi;
  const __half* f8_w_lo;
};
uint32_t tc_patch_a_offset(int dy, int dx, int row_bytes);
// a_hi / a_lo: plain activation views with box [Cin, 10, 1, 18, 1]; row_bytes = 2*Cin (128 or 64)
cudaError_t launch_tc_conv_patch(int row_bytes, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                 const CUtensorMap& w_lo, const TcPatchParams& prm, int num_sms, int* error_flag,
                                 cudaStream_t stream);

// ---- colour + conv1 on the tensor cores (tc_conv1.cu) ----------------------------------------------
struct TcConv1Params {
  const uint8_t* rgb;         // u8 [N,H,W,3] (colour transform fused), or
  const float* planes;        // f32 [3N,H,W,1]
  int N, H, W, Ho, Wo, pad_t, pad_l;
  unsigned long long wait_timeout;   // bound of a barrier wait in SM cycles, 0 = unbounded
  int Hs, Ws;                 // storage rows / columns per plane of the output (>= Ho, Wo)
  const __half* w_hi;         // device [2 sets][32 channels][32 taps (25 used)] fp16, scaled
  const __half* w_lo;
  const float* bias;          // device [2][32]
  float inv_scale[2];
  ColourConsts cc;
  __half* out_hi;             // split fp16 [3N,Ho,Wo,32] ...
  __half* out_lo;
  float* out_f32;             // ... or fp32
};
cudaError_t launch_tc_conv1(const TcConv1Params& prm, int num_sms, int* error_flag, cudaStream_t stream);

// ---- dconv8 + colour inverse + pack on the tensor cores (tc_dconv8.cu) ------------------------------
struct TcDconv8Params {
  int N, Hi, Wi;              // images, input size (output is 2Hi x 2Wi)
  int fast;                   // activations as one fp16 plane: A_hi x [W_hi | W_lo] only
  unsigned long long wait_timeout;   // bound of a barrier wait in SM cycles, 0 = unbounded
  float inv_scale[2];         // 2^-(ka+kw) per weight set
  float bias[2];
  ColourConsts cc;
  uint8_t* rgb;               // optional u8 [N,2Hi,2Wi,3]
  float* prequant;            // optional f32 [N,2Hi,2Wi,3]
  float* planes_out;          // optional f32 [3][N,2Hi,2Wi,1]
};
// a_hi / a_lo: plain views of the split-fp16 input [3N,Hi,Wi,64] with box [64, 8, 1, 16, 1];
// w_hi / w_lo: [2 sets x 32 taps (25 used)][64] fp16 matrices with box [64, 32]
cudaError_t launch_tc_dconv8(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                             const CUtensorMap& w_lo, const TcDconv8Params& prm, int num_sms, int* error_flag,
                             cudaStream_t stream);

// dconv8 behind the fused dconv7 (TcPatchParams::f8_out): R [3N][tile][25][4][128] fp32 -> rgb / prequant / planes_out as above,
// prm.Hi / prm.Wi = 2Hp x 2Wp.
cudaError_t launch_dconv8_gather(const float* R, const TcDconv8Params& prm, int Hp, int Wp, cudaStream_t stream);

}  // namespace nnic
