// Round-2 hardware probes (run on a B200) for the CTA-pair form of the convolution kernel:
//   E. tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, each CTA supplies its 128 rows of A and HALF of the B rows):
//      numerical check of the operand placement, then the issue rate per k-step for the split pattern
//      (A_hi x [W_hi|W_lo] wide + A_lo x W_hi narrow) in 1-CTA and 2-CTA form;
//   F. the same rates while a second warp streams weight-tile-sized bulk copies into shared memory at the rate the
//      convolution kernel does (tests the model "the kernel is bound by the shared-memory port: MMA operand reads +
//      TMA fills").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tc2 tools/probe_tc2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                          \
  do {                                                                                    \
    cudaError_t e = (x);                                                                  \
    if (e != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
      exit(2);                                                                            \
    }                                                                                     \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const unsigned long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if ((unsigned long long)clock64() - t0 > 2000000000ull) __trap();
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// arrives on the barrier at this shared-memory offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CG, int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
}
template <int CG, int COLS>
__device__ __forceinline__ void tmem_free(uint32_t addr) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS));
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS));
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ inline uint32_t sw128_off(int row, int k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2; }

template <int CG>
static cudaError_t launch_cg(void (*kern)(const uint8_t*, int, int, float*), int pairs, int threads, int smem, const uint8_t* a, int b, int c, float* d) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * CG); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a, b, c, d);
}

// ---------------------------------------------------------------------------------------------
// Probe E1: one M = 256 MMA chain over a CTA pair.  Per CTA the shared-memory image is
//   [A: 128 rows x 64 fp16, SWIZZLE_128B (16 KB)] [B half: N/2 rows x 64 fp16 (up to 16 KB)]
// out[cta][128][N] = the CTA's TMEM lanes.  Expected: out[cta][m][n] = sum_k A_cta[m][k] * B[n][k], where B rows
// [0, N/2) come from CTA 0 and rows [N/2, N) from CTA 1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pair_check(const uint8_t* __restrict__ images, int image_bytes, int N, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = cluster_ctarank();
  const uint8_t* image = images + (size_t)rank * image_bytes;
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc<2, 256>(&tmem_slot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                         // both CTAs' operands and barriers are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(256, N);
    const uint32_t base = smem_u32(smem);
    for (uint32_t ks = 0; ks < 4; ++ks)
      umma<2>(tmem, make_desc(base, 1024) + ks * 2, make_desc(base + 16384, 1024) + ks * 2, idesc, ks ? 1u : 0u);
    umma_commit<2>(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) out[((size_t)rank * 128 + m) * N + c + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_free<2, 256>(tmem);
}

static float frand(uint32_t& s) { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xffff) / 65536.0f; }

static void probe_e1() {
  printf("== probe E1: tcgen05.mma.cta_group::2, M = 256 over a CTA pair: operand placement ==\n");
  for (int N : {128, 64, 256}) {
    const int IMG = 16384 + 16384;
    std::vector<uint8_t> img(2 * IMG, 0);
    std::vector<float> A(256 * 64), B(N * 64);
    uint32_t seed = 3;
    for (int r = 0; r < 256; ++r) for (int k = 0; k < 64; ++k) {
      __half h = __float2half_rn(frand(seed) * 2 - 1); A[r * 64 + k] = __half2float(h);
      *reinterpret_cast<__half*>(&img[(r / 128) * IMG + sw128_off(r % 128, k)]) = h;
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) {
      __half h = __float2half_rn(frand(seed) * 2 - 1); B[n * 64 + k] = __half2float(h);
      *reinterpret_cast<__half*>(&img[(n / (N / 2)) * IMG + 16384 + sw128_off(n % (N / 2), k)]) = h;
    }
    uint8_t* d_img; float* d_out;
    CHECK(cudaMalloc(&d_img, img.size()));
    CHECK(cudaMalloc(&d_out, 256 * N * 4));
    CHECK(cudaMemset(d_out, 0, 256 * N * 4));
    CHECK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    const int smem = IMG + 1024;
    CHECK(cudaFuncSetAttribute(k_pair_check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CHECK(launch_cg<2>(k_pair_check, 1, 128, smem, d_img, IMG, N, d_out));
    CHECK(cudaDeviceSynchronize());
    std::vector<float> got(256 * N);
    CHECK(cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost));
    double maxd = 0;
    for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)A[m * 64 + k] * B[n * 64 + k];
      maxd = fmax(maxd, fabs(s - got[m * N + n]));
    }
    printf("  N = %3d (each CTA holds %3d rows of B): max|diff| = %.3e  %s\n", N, N / 2, maxd, maxd < 1e-3 ? "OK" : "WRONG");
    cudaFree(d_img); cudaFree(d_out);
  }
}

// ---------------------------------------------------------------------------------------------
// Probe E2 / F: issue rate.  Thread 0 of the (leader) CTA issues, per k-step, a wide MMA (N = nw) and optionally a
// narrow one (N = nn) -- the split pattern -- over 4 k-steps per group, alternating between two A stages.
// Warp 1 of EVERY CTA streams `fill_bytes` bulk copies (global, L2-resident -> shared) paced at one copy per
// `pace` cycles (0 = no fills), four in flight, like the weight-tile ring of the convolution kernel.
// ---------------------------------------------------------------------------------------------
struct RateArgs { int nw, nn, iters, fill_bytes, pace; };

template <int CG>
__global__ void __launch_bounds__(128) k_rate(const uint8_t* __restrict__ src, RateArgs ra, long long* __restrict__ cycles_out,
                                              int* __restrict__ fills_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, fbar[4];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  constexpr int A_BYTES = 4 * 16384, W_BYTES = 2 * 32768, RING = 4 * 16384;
  for (int i = threadIdx.x * 16; i < A_BYTES + W_BYTES + RING; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&fbar[i], 1);
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc<CG, 512>(&tmem_slot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    if (rank == 0) {
      const uint32_t idw = make_idesc(128 * CG, ra.nw), idn = make_idesc(128 * CG, ra.nn);
      const uint32_t base = smem_u32(smem);
      uint64_t a[4], w[2];
      for (int i = 0; i < 4; ++i) a[i] = make_desc(base + i * 16384, 1024);
      for (int i = 0; i < 2; ++i) w[i] = make_desc(base + A_BYTES + i * 32768, 1024);
      const long long t0 = clock64();
      for (int it = 0; it < ra.iters; ++it) {
        const int st = it & 1;
        const uint64_t ah = a[2 * st], al = a[2 * st + 1], wh = w[st];
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma<CG>(tmem, ah + ks * 2, wh + ks * 2, idw, 1u);
          if (ra.nn) umma<CG>(tmem + 256, al + ks * 2, wh + ks * 2, idn, 1u);
        }
      }
      umma_commit<CG>(&bar);
      mbar_wait(&bar, 0);
      cycles_out[blockIdx.x / CG] = clock64() - t0;
    } else {
      mbar_wait(&bar, 0);
    }
    done = 1;
  } else if (threadIdx.x == 32 && ra.pace > 0) {
    // paced bulk copies into a 4-slot ring
    int issued = 0;
    size_t off = ((size_t)blockIdx.x * 7919 * 16384) % ((size_t)30 << 20);
    off &= ~(size_t)1023;
    long long next = clock64();
    while (!done) {
      if (clock64() < next) continue;
      next += ra.pace;
      const int s = issued & 3;
      if (issued >= 4) mbar_wait(&fbar[s], ((issued - 4) >> 2) & 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fbar[s])), "r"(ra.fill_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + A_BYTES + W_BYTES + s * 16384)), "l"(src + off), "r"(ra.fill_bytes), "r"(smem_u32(&fbar[s])) : "memory");
      off += 16384 * 3;
      if (off + 16384 > ((size_t)31 << 20)) off = 0;
      ++issued;
    }
    // drain
    for (int k = issued < 4 ? 0 : issued - 4; k < issued; ++k) mbar_wait(&fbar[k & 3], (k >> 2) & 1);
    fills_out[blockIdx.x] = issued;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (threadIdx.x < 32) tmem_free<CG, 512>(tmem);
}

template <int CG>
static void run_rate(int num_sms, const uint8_t* d_src, long long* d_cyc, int* d_fills, RateArgs ra, const char* label) {
  const int smem = 4 * 16384 + 2 * 32768 + 4 * 16384 + 1024;
  CHECK(cudaFuncSetAttribute(k_rate<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int pairs = num_sms / CG;
  CHECK(cudaMemset(d_fills, 0, num_sms * sizeof(int)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CHECK(cudaLaunchKernelEx(&cfg, k_rate<CG>, d_src, ra, d_cyc, d_fills));
  CHECK(cudaDeviceSynchronize());
  std::vector<long long> cyc(pairs); std::vector<int> fills(num_sms);
  CHECK(cudaMemcpy(cyc.data(), d_cyc, pairs * sizeof(long long), cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(fills.data(), d_fills, num_sms * sizeof(int), cudaMemcpyDeviceToHost));
  long long mx = 0; for (long long c : cyc) mx = c > mx ? c : mx;
  long long nf = 0; for (int f : fills) nf += f;
  const double per = (double)mx / (ra.iters * 4.0);
  const double ideal = (ra.nw + ra.nn) / 2.0;
  const double fill_rate = (double)nf / (pairs * CG) * ra.fill_bytes / (double)mx;
  printf("  cta_group::%d %-28s wide N %3d narrow N %3d fills %5d B / %4d cyc: %6.1f cycles per k-step (math %5.1f) -> pipe %4.1f%%, fill %.1f B/clk/SM\n",
         CG, label, ra.nw, ra.nn, ra.pace ? ra.fill_bytes : 0, ra.pace, per, ideal, 100.0 * ideal / per, fill_rate);
}

static void probe_e2f(int num_sms) {
  printf("== probe E2/F: MMA issue rate, split pattern, 1-CTA vs CTA pair, with and without concurrent smem fills (%d SMs) ==\n", num_sms);
  long long* d_cyc; int* d_fills; uint8_t* d_src;
  CHECK(cudaMalloc(&d_cyc, num_sms * sizeof(long long)));
  CHECK(cudaMalloc(&d_fills, num_sms * sizeof(int)));
  CHECK(cudaMalloc(&d_src, (size_t)32 << 20));
  CHECK(cudaMemset(d_src, 0, (size_t)32 << 20));
  const int it = 4096;
  // no fills
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {128, 0, it, 0, 0}, "wide only");
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {64, 0, it, 0, 0}, "narrow only");
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 0, 0}, "split (today)");
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {256, 128, it, 0, 0}, "split, 2 phases stacked");
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {64, 32, it, 0, 0}, "split, COUT 32 (conv8)");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {128, 0, it, 0, 0}, "wide only");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {64, 0, it, 0, 0}, "narrow only");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {32, 0, it, 0, 0}, "N 32 only");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {256, 0, it, 0, 0}, "N 256 only");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 0, 0}, "split");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {256, 128, it, 0, 0}, "split, 2 phases stacked");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {64, 32, it, 0, 0}, "split, COUT 32 (conv8)");
  // with fills: one weight tile per tap = 4 k-steps.  1-CTA: 16 KB per ~450-580 cycles; pair: 8 KB per CTA.
  for (int pace : {600, 520, 448}) {
    run_rate<1>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 16384, pace}, "split + W fills");
    run_rate<2>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 8192, pace}, "split + W fills (half)");
  }
  run_rate<1>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 16384, 300}, "split + heavy fills");
  run_rate<2>(num_sms, d_src, d_cyc, d_fills, {128, 64, it, 16384, 448}, "split + full-size fills");
  cudaFree(d_cyc); cudaFree(d_fills); cudaFree(d_src);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  const char* which = argc > 1 ? argv[1] : "ef";
  for (const char* p = which; *p; ++p) {
    if (*p == 'e') probe_e1();
    if (*p == 'f') probe_e2f(prop.multiProcessorCount);
    fflush(stdout);
  }
  return 0;
}
