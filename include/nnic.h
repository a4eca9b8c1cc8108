/*
 * nnic.h -- C ABI of libnnic.so, the B200 (sm_100a) implementation of the inference hot path of
 * AlexFuster/Neural_network_image_compression's tf2_0 codec.
 *
 * The reference has no FFI layer: its boundary for this path is the Python call surface of
 * tf2_0/src/encoder.py, decoder.py and utils.py.  Every entry point below names the reference
 * call it replaces (paths relative to the reference root).  Plain pointers and sizes only; no
 * C++ exception crosses this boundary; every function returns 0 on success or a negative
 * nnic_status, and nnic_last_error() gives the text for the most recent failure on a handle.
 *
 * Threading: a handle is bound to one CUDA device and is not thread-safe.  One handle per GPU.
 * Calls on one handle share its scratch memory and weights: with NNIC_MEM_DEVICE they must be stream-ordered
 * (same stream, or the caller synchronises between streams).
 * Memory kinds: NNIC_MEM_HOST pointers are ordinary host memory (pageable or pinned); the call
 * copies in/out and returns after the result is in the caller's buffer.  NNIC_MEM_DEVICE pointers
 * are device memory on the handle's device; the call only enqueues work on `stream` and returns.
 */
#ifndef NNIC_H_
#define NNIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

typedef struct nnic_handle nnic_t;

enum nnic_status {
  NNIC_OK = 0,
  NNIC_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, bad enum */
  NNIC_ERR_SHAPE = -2,         /* shape not supported by a development toggle (never returned by a default handle) */
  NNIC_ERR_CUDA = -3,          /* a CUDA runtime/driver call failed; see nnic_last_error */
  NNIC_ERR_NO_WEIGHTS = -4,    /* a network needed by the call has unset layers */
  NNIC_ERR_NO_DEVICE = -5      /* no usable sm_100 device */
};

enum nnic_mem_kind { NNIC_MEM_HOST = 0, NNIC_MEM_DEVICE = 1 };

/* The four networks (tf2_0/src/utils.py:15-28: ProClass.models[0] = 'Y', models[1] = 'CbCr',
 * one ProClass for Encoder and one for Decoder). */
enum nnic_weight_set { NNIC_SET_ENC_Y = 0, NNIC_SET_ENC_CBCR = 1, NNIC_SET_DEC_Y = 2, NNIC_SET_DEC_CBCR = 3 };

/* Layer index inside a network, in call order (encoder.py:10-17: conv1, conv2, conv3, conv4, conv8;
 * decoder.py:10-17: dconv1, dconv5, dconv6, dconv7, dconv8). */
enum { NNIC_LAYERS_PER_NET = 5 };

/* Arithmetic of the eight GEMM-shaped layers (conv2/3/4/8, dconv1/5/6/7):
 *   NNIC_ARITH_TC_SPLIT  tcgen05 tensor cores, fp16 hi+lo split operands (3 MMAs per product),
 *                        fp32 accumulation in TMEM.  Default.  Any image size >= 1x1 (odd sizes at any stage are
 *                        stored with an even, zero-filled pitch so that the stride-2 parity views stay valid).
 *   NNIC_ARITH_SIMT_F32  plain fp32 FFMA kernels (any H, W >= 1).  Cross-check path.
 * conv1, dconv8, colour, quantise, histogram and pack kernels are fp32/integer in both modes. */
enum nnic_arith { NNIC_ARITH_TC_SPLIT = 0, NNIC_ARITH_SIMT_F32 = 1 };

/* Optional reduced precision of the DECODER under NNIC_ARITH_TC_SPLIT (the encoder is never affected: its symbols
 * must be bit-exact).  BASELINE.json asks reconstructions to be within 0.01 dB PSNR of the reference, not bit-exact:
 *   NNIC_DECODE_SPLIT  default: as exact as the encoder (0 differing bytes against the fp32 restatement in the tests)
 *   NNIC_DECODE_FP16   dconv1..dconv7: ONE fp16 product per MAC (activations and weights rounded to fp16, fp32
 *                      accumulation in TMEM, activations stored as one fp16 plane); dconv8: A_hi x (W_hi + W_lo).
 *                      About 1/3 of the tensor work and 1/2 of the activation traffic; reconstruction bytes differ
 *                      from the exact decoder by +-1 in a few per cent of the positions (tests state the PSNR bound). */
enum nnic_decode_precision { NNIC_DECODE_SPLIT = 0, NNIC_DECODE_FP16 = 1 };

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Replaces: constructing Encoder()/Decoder() (encoder.py:34-36, decoder.py:35-37). */
int nnic_create(int device, nnic_t** out);
void nnic_destroy(nnic_t* h);
const char* nnic_last_error(const nnic_t* h);   /* never NULL; h may be NULL (global message) */
const char* nnic_version(void);
int nnic_set_arith(nnic_t* h, int arith);
int nnic_get_arith(const nnic_t* h);
int nnic_set_decode_precision(nnic_t* h, int precision);
int nnic_get_decode_precision(const nnic_t* h);
/* Number of kernel launches this handle has enqueued since creation (bench.py's gpu_launches). */
uint64_t nnic_launch_count(const nnic_t* h);

/* ---- weights --------------------------------------------------------------------------------
 * Replaces: ProClass.load -> Keras load_weights (utils.py:26-28).  `kernel` is in the Keras layout
 * of the layer (Conv2D [kh,kw,Cin,Cout]; Conv2DTranspose [kh,kw,Cout,Cin]), fp32, host memory;
 * `bias` is [Cout].  The library keeps its own repacked copies. */
int nnic_set_weights(nnic_t* h, int set, int layer, const float* kernel, const float* bias);

/* Replaces: a freshly constructed Encoder() / Decoder() (encoder.py:34-36, decoder.py:35-37), whose Keras layers
 * start from glorot-uniform kernels (limit sqrt(6 / ((Cin + Cout) * kh * kw))) and zero biases.
 * nnic_init_random installs such a network as weight set `set`.  The stream is NumPy's default_rng(seed) (PCG64 seeded
 * through SeedSequence), i.e. bit-identical to neural_network_image_compression_b200/weights.py::glorot_uniform(kind, seed)
 * -- the weight sets the oracle and the golden vectors use (seeds 11, 12, 13, 14 for the four sets).
 * nnic_init_random_scaled multiplies the kernels by `gain` and draws biases from U(-bias_range, bias_range)
 * (bias_range = 0: zero biases): the "spread" sets of the parity tests are (1.6, 0.05).
 * nnic_glorot_uniform fills caller buffers instead (no handle, no GPU): the five kernels back to back in their Keras
 * layouts, the five bias vectors back to back; nnic_layer_shape gives k, Cin, Cout of a layer. */
int nnic_init_random(nnic_t* h, int set, uint64_t seed);
int nnic_init_random_scaled(nnic_t* h, int set, uint64_t seed, double gain, double bias_range);
int nnic_glorot_uniform(int set, uint64_t seed, double gain, double bias_range, float* kernels, float* biases);
int nnic_layer_shape(int set, int layer, int* ksize, int* cin, int* cout);

/* ---- encode ---------------------------------------------------------------------------------
 * Replaces: Encoder.__call__ (encoder.py:38-47).
 *   rgb     uint8 [N,H,W,3], C-contiguous
 *   latent  uint8 [N,ceil(H/8),ceil(W/8),96]; channels 0-31 Y, 32-63 Cb, 64-95 Cr
 *   prequant optional (may be NULL) float [N,h,w,96]: the clipped encoder output before
 *            *255/round (what tf.concat(encoded) holds at encoder.py:45) */
int nnic_encode(nnic_t* h, const uint8_t* rgb, int N, int H, int W, uint8_t* latent, float* prequant,
                int mem_kind, void* stream);

/* ---- decode ---------------------------------------------------------------------------------
 * Replaces: Decoder.__call__ (decoder.py:39-48).
 *   latent  uint8 [N,lh,lw,96]
 *   rgb     uint8 [N,8*lh,8*lw,3]
 *   prequant optional (may be NULL) float [N,8*lh,8*lw,3]: clipped RGB in [0,1] before *255/round */
int nnic_decode(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, uint8_t* rgb, float* prequant,
                int mem_kind, void* stream);

/* ---- plane-level model calls ----------------------------------------------------------------
 * Replaces: ProClass.run_model (utils.py:19-24): three single-network calls with weight sets (0,1,1).
 *   encoder: planes float [3][N,H,W,1] (plane-major) -> out float [3][N,h,w,32], clipped to [0,1]
 *   decoder: planes float [3][N,lh,lw,32]            -> out float [3][N,8lh,8lw,1], clipped to [0,1] */
int nnic_run_encoder_planes(nnic_t* h, const float* planes, int N, int H, int W, float* out,
                            int mem_kind, void* stream);
int nnic_run_decoder_planes(nnic_t* h, const float* planes, int N, int lh, int lw, float* out,
                            int mem_kind, void* stream);

/* ---- rate -----------------------------------------------------------------------------------
 * Replaces: the discrete histogram/entropy block of tf1_13/src/training.py:62-71 (the only
 * histogram entropy in the reference), applied to the uint8 latent of Encoder.__call__.
 *   latent        uint8 [N,lh,lw,96]
 *   H, W          size of the source images (for bits per pixel)
 *   hist          optional uint32 [N][3][256]   per (image, plane) counts
 *   entropy_bits  optional float  [N][3]        sum p*(-log(clip(p,1e-5,1))/log 2), bits per symbol
 *   bpp           optional float  [N]           sum_p entropy*(lh*lw*32)/(H*W)   (this build's definition)
 *   hist_global   optional uint64 [3][256]      counts summed over the N images; ACCUMULATED into
 *                                               (caller zeroes it), so calls over micro-batches add up */
int nnic_rate(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, int H, int W, uint32_t* hist,
              float* entropy_bits, float* bpp, uint64_t* hist_global, int mem_kind, void* stream);

/* Encode and rate in one pass: Encoder.__call__ followed by the histogram block above, with the symbol counts
 * taken inside the kernel that quantises the latent (conv8's epilogue: shared-memory histogram of the tile,
 * flushed with one global atomic per non-empty bin), so the latent is not read again.  Outputs as in nnic_encode
 * and nnic_rate; results are identical to calling the two one after the other. */
int nnic_encode_rate(nnic_t* h, const uint8_t* rgb, int N, int H, int W, uint8_t* latent, uint32_t* hist,
                     float* entropy_bits, float* bpp, uint64_t* hist_global, int mem_kind, void* stream);

/* Optional finer table: symbol counts per latent FEATURE CHANNEL, summed over the N images (BASELINE.json north_star
 * "per-channel latent histogram"; the reference's own histogram is the per-(image, plane) one of nnic_rate, and rows
 * 32p .. 32p+31 of this table add up to hist_global[p]).
 *   hist_channels  uint64 [96][256], ACCUMULATED into (caller zeroes it) */
int nnic_rate_channels(nnic_t* h, const uint8_t* latent, int N, int lh, int lw, uint64_t* hist_channels, int mem_kind,
                       void* stream);

/* Cross-rank sum of the global symbol counts: the path's only exchange step.
 *   nccl_comm    the caller's ncclComm_t for this rank (one rank per GPU)
 *   hist_global  DEVICE uint64 [3][256], summed in place over all ranks of the communicator
 * Enqueues one ncclAllReduce(sum, uint64, 768) on `stream`.  NCCL is resolved at call time from the library the
 * process has already loaded (so the communicator and the call come from the same NCCL), else from libnccl.so.2;
 * libnnic.so has no link-time dependency on it.  Integer sums: the result is identical for any number of ranks. */
int nnic_hist_allreduce(nnic_t* h, void* nccl_comm, uint64_t* hist_global, void* stream);

/* Entropy of already-reduced counts (e.g. hist_global after the cross-rank allreduce):
 *   counts uint64 [rows][256] -> entropy_bits float [rows].  Same formula as nnic_rate. */
int nnic_entropy_from_counts(nnic_t* h, const uint64_t* counts, int rows, float* entropy_bits,
                             int mem_kind, void* stream);

/* ---- forward extras of the training step ----------------------------------------------------
 * The reference's training loop runs, beside encoder and decoder, three forward computations (tf2_0/src/training.py):
 *
 * nnic_entropynet_*: Entropynet (training.py:25-42), the CNN regressor of the PNG rate: Conv2D(64,5,2,SAME,leaky),
 *   2 x Conv2D(64,3,1,SAME,leaky), Flatten, Dense(512), Dense(1), clip(0, 8), applied to the encoder's float output
 *   `batch_encoded` [P,h,w,32] (P = 3N planes; ONE network for all planes).  Layers 0..4 = conv1, conv2, conv3 (Keras Conv2D
 *   kernels [kh,kw,Cin,Cout]), dense1 (Keras Dense kernel [features,512], features = 64*ceil(h/2)*ceil(w/2), NHWC flatten order),
 *   dense2 ([512,1]); `features` is only read for layer 3.  The convolutions run on the tensor-core convolution kernel
 *   (FFMA kernels for odd h or w), the dense layers in fp32.  approx_entropy: float [P].
 * nnic_noise_quantise: the quantisation proxy (training.py:87-88) out = clip(encoded + u/255, 0, 1), u ~ U(-0.5, 0.5).
 *   `noise` (optional, `count` floats) supplies u; otherwise u is drawn from Philox4x32-10 keyed by `seed` and the element
 *   index, so the result does not depend on the launch geometry or on the rank count.
 * nnic_ssim: tf.image.ssim(a, b, max_val=1.0) (training.py:108,113) of single-channel float images [P,H,W] (H, W >= 11):
 *   11x11 Gaussian window (sigma 1.5), k1 = 0.01, k2 = 0.03, mean over the (H-10) x (W-10) map.  ssim: float [P]. */
int nnic_entropynet_set_weights(nnic_t* h, int layer, const float* kernel, const float* bias, int features);
int nnic_entropynet_forward(nnic_t* h, const float* encoded, int P, int lh, int lw, float* approx_entropy, int mem_kind,
                            void* stream);
int nnic_noise_quantise(nnic_t* h, const float* encoded, size_t count, uint64_t seed, const float* noise, float* out,
                        int mem_kind, void* stream);
int nnic_ssim(nnic_t* h, const float* a, const float* b, int P, int H, int W, float* ssim, int mem_kind, void* stream);

/* ---- scratch management ---------------------------------------------------------------------
 * Activation scratch is grown lazily and reused.  Images beyond `max_planes_in_flight` colour
 * planes are processed in micro-batches inside one call.  0 = library default. */
int nnic_set_micro_batch(nnic_t* h, int max_images_in_flight);
/* CUtensorMap encodes this handle has done since creation (activation views are cached per tensor and shape: a
 * steady-state call encodes none). */
uint64_t nnic_tensor_map_encodes(const nnic_t* h);
size_t nnic_scratch_bytes(const nnic_t* h);

/* ---- per-kernel timing -------------------------------------------------------------------------
 * With profiling on, every kernel launch is bracketed by CUDA events on the launching stream.
 * nnic_profile_collect waits for them, writes the summed milliseconds and the launch count per kernel id
 * (NNIC_KERNEL_*; arrays of at least NNIC_KERNEL_COUNT entries) since the previous collect, and returns
 * NNIC_KERNEL_COUNT.  Used by bench.py for the roofline figures. */
enum nnic_kernel_id {
  NNIC_KERNEL_CONV1 = 0, NNIC_KERNEL_CONV2, NNIC_KERNEL_CONV3, NNIC_KERNEL_CONV4, NNIC_KERNEL_CONV8,
  NNIC_KERNEL_QUANTISE, NNIC_KERNEL_EXPAND, NNIC_KERNEL_DCONV1, NNIC_KERNEL_DCONV5, NNIC_KERNEL_DCONV6,
  NNIC_KERNEL_DCONV7, NNIC_KERNEL_DCONV8, NNIC_KERNEL_HIST, NNIC_KERNEL_ENTROPY, NNIC_KERNEL_HIST_REDUCE,
  NNIC_KERNEL_F32_SPLIT, NNIC_KERNEL_ENTROPYNET_CONV, NNIC_KERNEL_DENSE, NNIC_KERNEL_SSIM, NNIC_KERNEL_NOISE, NNIC_KERNEL_COUNT
};
int nnic_set_profiling(nnic_t* h, int on);
int nnic_profile_collect(nnic_t* h, float* ms_per_kernel, int* launches_per_kernel, int capacity);

/* ---- introspection used by the parity tests ----------------------------------------------------
 * nnic_colour_constants: the fp32 colour matrices the kernels use: (float32)ycbcr_kernel,
 * (float32)np.linalg.inv(ycbcr_kernel), (float32)ycbcr_off (utils.py:7-9).  Any pointer may be NULL.
 * nnic_debug_fetch: copy an intermediate activation of the most recent encode (slots 0-3: conv1, conv2,
 * conv3, conv4+res) or decode (slots 4-7: latent/255, dconv1, dconv5, dconv6+res) micro-batch to host
 * memory as fp32 [3*nb,H,W,C].  out == NULL returns the element count. */
void nnic_colour_constants(float* k9, float* kinv9, float* off3);
long long nnic_debug_fetch(nnic_t* h, int slot, float* out, long long capacity);
/* The tensor-core layers keep activations as two fp16 planes of v*16, which saturate at |v| > 4094 where the fp32
 * reference would carry on.  nnic_debug_saturated counts the saturated values in the activations of the most recent
 * encode and decode micro-batch (synchronises the device): 0 = the representation was exact to its ~22 bits.  The
 * decoder's last activation (dconv7's output) exists on chip only in the default fused form and is counted when the
 * handle was created with NNIC_FUSE_D78=0. */
long long nnic_debug_saturated(nnic_t* h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NNIC_H_ */
