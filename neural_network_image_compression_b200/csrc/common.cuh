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
