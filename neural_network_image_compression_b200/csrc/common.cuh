// Shared definitions for the libnnic kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nnic {

// ---- activation storage between tensor-core layers -------------------------------------------
// A value v is kept as two fp16 planes: hi = fp16(v*ACT_SCALE), lo = fp16(v*ACT_SCALE - hi).
// hi+lo carries ~22 significant bits; ACT_SCALE (a power of two, exact) keeps lo out of the fp16
// subnormal range for |v| >= 2^-7.  |v*ACT_SCALE| is saturated to the fp16 maximum.
constexpr float ACT_SCALE = 16.0f;
constexpr float ACT_INV_SCALE = 1.0f / 16.0f;
constexpr float FP16_MAX = 65504.0f;
constexpr float LEAKY_ALPHA = 0.2f;   // tf.nn.leaky_relu default (reference encoder.py:10-17)

__device__ __forceinline__ void split_f32(float v, __half& hi, __half& lo) {
  float vs = fminf(fmaxf(v * ACT_SCALE, -FP16_MAX), FP16_MAX);
  hi = __float2half_rn(vs);
  lo = __float2half_rn(vs - __half2float(hi));
}
// Two values at once: {hi(v0), hi(v1)} and {lo(v0), lo(v1)} as packed fp16 pairs (v0 in the low half).  The
// saturating pack instruction (F2FP.SATFINITE) replaces the explicit clamp of split_f32.
// split2_scaled takes values that already carry the factor ACT_SCALE.
__device__ __forceinline__ void split2_scaled(float s0, float s1, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(s1), "f"(s0));
  const float f0 = __half2float(__ushort_as_half((unsigned short)(hi2 & 0xffffu)));
  const float f1 = __half2float(__ushort_as_half((unsigned short)(hi2 >> 16)));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(s1 - f1), "f"(s0 - f0));
}
// ---- packed fp32 pairs (FADD2 / FMUL2 / FFMA2 of sm_100): two IEEE round-to-nearest operations per instruction, bit-identical to
// the scalar forms; the epilogues are bound by instruction issue, not by the fp32 lanes ----
typedef unsigned long long f32x2_t;           // {low 32 bits: first value, high 32 bits: second value}
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2_t pack2u(uint32_t a, uint32_t b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// split of a pair that already carries the factor ACT_SCALE (see split2_scaled): one packed subtraction for the two residuals
__device__ __forceinline__ void split2_scaled_x2(f32x2_t s, uint32_t& hi2, uint32_t& lo2) {
  float s0, s1;
  unpack2(s, s0, s1);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(s1), "f"(s0));
  const float f0 = __half2float(__ushort_as_half((unsigned short)(hi2 & 0xffffu)));
  const float f1 = __half2float(__ushort_as_half((unsigned short)(hi2 >> 16)));
  float d0, d1;
  unpack2(sub2(s, pack2(f0, f1)), d0, d1);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(d1), "f"(d0));
}
// bias + leaky_relu of a pair: v = acc * scale + bias (one rounding: scale is a power of two), max(v, 0.2 v)
__device__ __forceinline__ f32x2_t bias_leaky2(f32x2_t acc, f32x2_t scale, f32x2_t bias) {
  const f32x2_t v = fma2(acc, scale, bias);
  float v0, v1, t0, t1;
  unpack2(v, v0, v1);
  unpack2(mul2(v, pack2(LEAKY_ALPHA, LEAKY_ALPHA)), t0, t1);
  return pack2(fmaxf(v0, t0), fmaxf(v1, t1));
}
__device__ __forceinline__ void split2_f32(float v0, float v1, uint32_t& hi2, uint32_t& lo2) {
  split2_scaled(v0 * ACT_SCALE, v1 * ACT_SCALE, hi2, lo2);
}
__device__ __forceinline__ float join_f32(__half hi, __half lo) {
  return (__half2float(hi) + __half2float(lo)) * ACT_INV_SCALE;
}
// programmatic dependent launch (kernels.h launch_kernel): let the next kernel of the stream start its prologue; wait until every
// earlier kernel of the stream has completed and its writes are visible
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// leaky_relu as TF computes it in fp32: one rounded multiply on the negative side.
__device__ __forceinline__ float leaky(float v) { return v > 0.0f ? v : __fmul_rn(v, LEAKY_ALPHA); }

// ---- tap programs ----------------------------------------------------------------------------
// Every GEMM-shaped layer is a set of "jobs" (one per output parity phase); a job is a list of
// steps, each step one tap (or tap pair) of the kernel:
//   out[(Y*out_s + out_oy), (X*out_s + out_ox), :] = sum_steps  in_view[Y + dy, X + dx, py, koff..] x Wstep
constexpr int MAX_STEPS = 25;
constexpr int MAX_JOBS = 4;

struct TcStep {
  int16_t dy, dx;      // offset of the tap in view rows / cols
  int16_t py;          // coordinate in the row-parity dimension of the view (0 for plain views)
  int16_t koff;        // element offset inside the view's innermost dimension
  int16_t w_row;       // first row of this step's [COUT x KSLAB] tile in the weight matrix
  int8_t ks_begin, ks_end;  // 16-element k-steps of the slab that carry non-zero weights
  int16_t pad_;
};
struct TcJob {             // host-side description; nnic_api.cu lowers it to the kernel's TcPatchJob
  int nsteps;
  int out_oy, out_ox;  // output offset of this phase
  TcStep steps[MAX_STEPS];
};

struct SimtTap { int16_t dy, dx, widx, pad_; };
struct SimtJob {
  int ntaps;
  int out_oy, out_ox;
  SimtTap taps[MAX_STEPS];
};
struct SimtJobs {
  int njobs;
  int in_stride;    // input sampling stride (2 for strided convs)
  int out_stride;   // output stride (2 for transposed convs)
  SimtJob job[MAX_JOBS];
};

}  // namespace nnic
