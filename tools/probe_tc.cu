// Hardware probes that decide the design of the tensor-core convolution (run on a B200):
//   A. does a UMMA shared-memory descriptor whose start address is shifted by whole 128-byte rows
//      inside a SWIZZLE_128B tile (and whose 8-row-group stride is not 1024) read the rows one expects?
//      (needed to take the 9 taps of a 3x3 convolution out of ONE halo patch in shared memory)
//   B. how accurate is the fp32 accumulation in TMEM over K = 1600 with fp16 hi/lo split operands?
//   C. how many cycles does a 128xNx16 fp16 MMA take from shared memory for N = 64 / 128 / 256?
//   D. L2 -> shared memory bandwidth per SM with all SMs streaming (TMA-free: cp.async.bulk 1D)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tc tools/probe_tc.cu
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
template <int COLS>
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  return 0;
}

// element (row, k) of a K-major SWIZZLE_128B tile whose rows are 128 bytes (64 fp16), tile base 1024-aligned:
// byte offset = row*128 + ((k/8) ^ (row%8))*16 + (k%8)*2
__host__ __device__ inline uint32_t sw128_off(int row, int k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2; }

// ---------------------------------------------------------------------------------------------
// Probe A / B kernel: generic "MMA over an smem image".  The host supplies a complete shared-memory image
// (already swizzled) and a list of MMA issue records {a_off, b_off, sbo_a, base_offset_a, accumulate}.
// ---------------------------------------------------------------------------------------------
struct MmaRec { uint32_t a_off, b_off, sbo_a, base_off_a, acc, ksteps; };

__global__ void __launch_bounds__(128) k_mma_image(const uint8_t* __restrict__ image, int image_bytes,
                                                   const MmaRec* __restrict__ recs, int nrecs, int N, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
  if (threadIdx.x < 32) tmem_alloc<256>(&tmem_slot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t base = smem_u32(smem);
    for (int r = 0; r < nrecs; ++r) {
      const MmaRec rec = recs[r];
      for (uint32_t ks = 0; ks < rec.ksteps; ++ks) {
        const uint64_t a = make_desc(base + rec.a_off, rec.sbo_a, rec.base_off_a) + ks * 2;
        const uint64_t b = make_desc(base + rec.b_off, 1024, 0) + ks * 2;
        umma(tmem, a, b, idesc, (rec.acc || ks > 0) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) out[m * N + c + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

static std::vector<float> run_image(const std::vector<uint8_t>& image, const std::vector<MmaRec>& recs, int N) {
  uint8_t* d_img; MmaRec* d_rec; float* d_out;
  CHECK(cudaMalloc(&d_img, image.size()));
  CHECK(cudaMalloc(&d_rec, recs.size() * sizeof(MmaRec)));
  CHECK(cudaMalloc(&d_out, 128 * N * 4));
  CHECK(cudaMemcpy(d_img, image.data(), image.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(d_rec, recs.data(), recs.size() * sizeof(MmaRec), cudaMemcpyHostToDevice));
  const int smem = (int)image.size() + 1024;
  CHECK(cudaFuncSetAttribute(k_mma_image, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_mma_image<<<1, 128, smem>>>(d_img, (int)image.size(), d_rec, (int)recs.size(), N, d_out);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  std::vector<float> out(128 * N);
  CHECK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_img); cudaFree(d_rec); cudaFree(d_out);
  return out;
}

static float frand(uint32_t& s) { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xffff) / 65536.0f; }

static void probe_a() {
  printf("== probe A: row-shifted / strided UMMA descriptors in a SWIZZLE_128B tile ==\n");
  const int ROWS = 256, N = 64;
  // image: A rows [0,256) at offset 0 (32 KB), B rows [0,64) at offset 32768 (8 KB)
  std::vector<uint8_t> img(32768 + 8192, 0);
  std::vector<float> A(ROWS * 64), B(N * 64);
  uint32_t seed = 1;
  for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 64; ++k) {
    __half h = __float2half_rn(frand(seed) * 2 - 1); A[r * 64 + k] = __half2float(h);
    *reinterpret_cast<__half*>(&img[sw128_off(r, k)]) = h;
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) {
    __half h = __float2half_rn(frand(seed) * 2 - 1); B[n * 64 + k] = __half2float(h);
    *reinterpret_cast<__half*>(&img[32768 + sw128_off(n, k)]) = h;
  }
  auto expect = [&](int shift, int group_stride_rows, std::vector<float>& e) {
    e.assign(128 * N, 0.f);
    for (int m = 0; m < 128; ++m) {
      const int row = (m / 8) * group_stride_rows + (m % 8) + shift;
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < 64; ++k) s += (double)A[row * 64 + k] * B[n * 64 + k];
        e[m * N + n] = (float)s;
      }
    }
  };
  struct Case { int shift, gstride, base_off; };
  std::vector<Case> cases;
  for (int s = 0; s <= 9; ++s) { cases.push_back({s, 8, 0}); if (s % 8) cases.push_back({s, 8, s % 8}); }
  for (int s = 0; s <= 2; ++s) { cases.push_back({s, 10, 0}); }
  cases.push_back({1, 10, 1}); cases.push_back({0, 12, 0}); cases.push_back({3, 12, 0});
  for (const Case& c : cases) {
    std::vector<MmaRec> recs = {{(uint32_t)c.shift * 128, 32768, (uint32_t)c.gstride * 128, (uint32_t)c.base_off, 0, 4}};
    std::vector<float> got = run_image(img, recs, N), e;
    expect(c.shift, c.gstride, e);
    double maxd = 0;
    for (size_t i = 0; i < e.size(); ++i) maxd = fmax(maxd, fabs((double)got[i] - e[i]));
    printf("  row shift %d, 8-row-group stride %3d rows (SBO %4d B), base_offset field %d : max|diff| = %.3e  %s\n", c.shift,
           c.gstride, c.gstride * 128, c.base_off, maxd, maxd < 1e-3 ? "OK" : "WRONG");
  }
}

static void probe_b() {
  printf("== probe B: fp32 accumulation in TMEM, K = 1600 (100 k-steps), fp16 hi/lo split, N = 64 ==\n");
  // smem image: 25 slabs of [A_hi 16K | A_lo 16K | W_hi 8K | W_lo 8K] would be 1.2 MB; instead keep 4 slabs
  // (K = 256) and issue them 6.25x: use distinct data per slab only for the 4 slabs, re-used with accumulate.
  const int N = 64, SLABS = 4, REPS = 6;    // K = 4*64*6 = 1536
  const int SLAB = 49152;
  std::vector<uint8_t> img(SLABS * SLAB, 0);
  std::vector<float> a32(SLABS * 128 * 64), w32(SLABS * N * 64);
  std::vector<float> ahi(a32.size()), alo(a32.size()), whi(w32.size()), wlo(w32.size());
  uint32_t seed = 7;
  for (int s = 0; s < SLABS; ++s) {
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) {
      float v = frand(seed) < 0.5f ? frand(seed) * 2.0f * 16.0f : -frand(seed) * 0.4f * 16.0f;   // leaky-like activations * ACT_SCALE
      __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
      size_t i = ((size_t)s * 128 + r) * 64 + k;
      a32[i] = v; ahi[i] = __half2float(h); alo[i] = __half2float(l);
      *reinterpret_cast<__half*>(&img[s * SLAB + sw128_off(r, k)]) = h;
      *reinterpret_cast<__half*>(&img[s * SLAB + 16384 + sw128_off(r, k)]) = l;
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) {
      float v = (frand(seed) * 2 - 1) * 0.05f * 524288.0f;    // glorot-ish weights * 2^19 -> |w| < 26215
      __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
      size_t i = ((size_t)s * N + n) * 64 + k;
      w32[i] = v; whi[i] = __half2float(h); wlo[i] = __half2float(l);
      *reinterpret_cast<__half*>(&img[s * SLAB + 32768 + sw128_off(n, k)]) = h;
      *reinterpret_cast<__half*>(&img[s * SLAB + 40960 + sw128_off(n, k)]) = l;
    }
  }
  for (int variant = 0; variant < 2; ++variant) {
    // variant 0: hi*hi, lo*hi, hi*lo interleaved per k-step group (what the conv kernel does, per slab)
    // variant 1: all hi*hi first, then the two small terms at the end
    std::vector<MmaRec> recs;
    bool first = true;
    auto push = [&](int s, int a_sub, int w_sub) {
      recs.push_back({(uint32_t)(s * SLAB + a_sub * 16384), (uint32_t)(s * SLAB + 32768 + w_sub * 8192), 1024, 0, first ? 0u : 1u, 4});
      first = false;
    };
    if (variant == 0) {
      for (int rep = 0; rep < REPS; ++rep) for (int s = 0; s < SLABS; ++s) { push(s, 0, 0); push(s, 1, 0); push(s, 0, 1); }
    } else {
      for (int rep = 0; rep < REPS; ++rep) for (int s = 0; s < SLABS; ++s) push(s, 0, 0);
      for (int rep = 0; rep < REPS; ++rep) for (int s = 0; s < SLABS; ++s) { push(s, 1, 0); push(s, 0, 1); }
    }
    std::vector<float> got = run_image(img, recs, N);
    // references: exact (fp64) value of the fp32 operands' dot product; and of the three split products
    double max_rel_exact = 0, sum_rel = 0, max_rel_split = 0, sum_signed = 0, max_rel_f32seq = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double exact = 0, split = 0, mag = 0; float seq = 0.f;
      for (int s = 0; s < SLABS; ++s) for (int k = 0; k < 64; ++k) {
        size_t ia = ((size_t)s * 128 + m) * 64 + k, iw = ((size_t)s * N + n) * 64 + k;
        exact += (double)a32[ia] * w32[iw];
        split += (double)ahi[ia] * whi[iw] + (double)alo[ia] * whi[iw] + (double)ahi[ia] * wlo[iw];
        mag += fabs((double)a32[ia] * w32[iw]);
      }
      exact *= REPS; split *= REPS; mag *= REPS;
      for (int rep = 0; rep < REPS; ++rep) for (int s = 0; s < SLABS; ++s) for (int k = 0; k < 64; ++k) {
        size_t ia = ((size_t)s * 128 + m) * 64 + k, iw = ((size_t)s * N + n) * 64 + k;
        seq = fmaf(a32[ia], w32[iw], seq);
      }
      const double g = got[m * N + n];
      // errors relative to sum|products| (what an fp32 ulp of the running accumulator scales with)
      max_rel_exact = fmax(max_rel_exact, fabs(g - exact) / mag);
      max_rel_split = fmax(max_rel_split, fabs(g - split) / mag);
      max_rel_f32seq = fmax(max_rel_f32seq, fabs((double)seq - exact) / mag);
      sum_rel += fabs(g - exact) / mag;
      sum_signed += (g - exact) / mag;
    }
    printf("  variant %d: |tmem-exact|/sum|prod| max %.3e mean %.3e ; vs exact split-product sum max %.3e ; "
           "signed mean error / sum|products| %.3e ; (fp32 sequential-FMA reference: max %.3e) ; 2^-24 = %.3e\n",
           variant, max_rel_exact, sum_rel / (128 * N), max_rel_split, sum_signed / (128 * N), max_rel_f32seq, ldexp(1.0, -24));
  }
}

// ---------------------------------------------------------------------------------------------
// Probe C: MMA rate from shared memory with a tight issue loop.  One CTA per SM; a single thread issues
// groups of 4 k-step MMAs with descriptors precomputed in registers, then waits for the final commit.
// mode 0: every MMA has shape 128xNx16, A tile alternates between two smem buffers
// mode 1: the conv kernel's split pattern per k-step: 3 MMAs of N   (hi*hi, lo*hi, hi*lo)
// mode 2: N-stacked split pattern per k-step: one MMA of 2N (A_hi x [W_hi|W_lo]) + one of N (A_lo x W_hi)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128) k_mma_rate(int N, int iters, int mode, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x * 16; i < 4 * 16384 + 2 * 32768; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(N), idesc2 = make_idesc(2 * N);
    const uint32_t base = smem_u32(smem);
    // A_hi(0), A_lo(1) for stage 0; A_hi(2), A_lo(3) for stage 1; W at 4*16384 (hi, then lo contiguous) per stage 32 KB
    uint64_t a[4], w[2];
    for (int i = 0; i < 4; ++i) a[i] = make_desc(base + i * 16384, 1024, 0);
    for (int i = 0; i < 2; ++i) w[i] = make_desc(base + 4 * 16384 + i * 32768, 1024, 0);
    const uint32_t w_lo_off = (uint32_t)(N * 128) >> 4;   // W_lo tile directly behind the W_hi tile
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it & 1;
      const uint64_t ah = a[2 * st], al = a[2 * st + 1], wh = w[st];
      if (mode == 0) {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma_acc(tmem, ah + ks * 2, wh + ks * 2, idesc);
      } else if (mode == 1) {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma_acc(tmem, ah + ks * 2, wh + ks * 2, idesc);
          umma_acc(tmem, al + ks * 2, wh + ks * 2, idesc);
          umma_acc(tmem, ah + ks * 2, wh + w_lo_off + ks * 2, idesc);
        }
      } else {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma_acc(tmem, ah + ks * 2, wh + ks * 2, idesc2);        // N-stacked [W_hi | W_lo]: 2N rows of B
          umma_acc(tmem + 256, al + ks * 2, wh + ks * 2, idesc);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static void probe_c(int num_sms) {
  printf("== probe C: cycles per k-step group (tight issue loop), SS operands, %d CTAs ==\n", num_sms);
  long long* d_cyc;
  CHECK(cudaMalloc(&d_cyc, num_sms * sizeof(long long)));
  const int smem = 4 * 16384 + 2 * 32768 + 1024;
  CHECK(cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int grid : {1, num_sms}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int N : {32, 64, 128, 256}) {
        if (mode == 2 && N > 64) continue;
        if (mode == 1 && N > 128) continue;
        const int iters = 4096;
        k_mma_rate<<<grid, 128, smem>>>(N, iters, mode, d_cyc);
        CHECK(cudaGetLastError());
        CHECK(cudaDeviceSynchronize());
        std::vector<long long> cyc(grid);
        CHECK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0; for (long long c : cyc) mx = c > mx ? c : mx;
        const double per = (double)mx / (iters * 4.0);      // cycles per k-step (16 K elements)
        const double useful = mode == 0 ? 1.0 : 3.0;        // products of 128xNx16 per k-step
        printf("  grid %3d mode %d N %3d: %.1f cycles per k-step -> %.0f MAC/cycle/SM raw (peak 4096), ideal %.0f cycles\n", grid, mode,
               N, per, useful * 128.0 * N * 16 / per, useful * N / 2.0);
      }
    }
  }
  cudaFree(d_cyc);
}

// ---------------------------------------------------------------------------------------------
// Probe D: L2 -> shared memory streaming bandwidth, cp.async.bulk (1D TMA), all SMs, L2-resident source
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bulk_bw(const uint8_t* __restrict__ src, size_t src_bytes, int chunk, int iters,
                                                 long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long t0 = clock64();
    size_t off = ((size_t)blockIdx.x * 7919 * chunk) % (src_bytes - chunk);
    off &= ~(size_t)1023;
    // 4 chunks in flight
    for (int it = 0; it < iters + 4; ++it) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(&bar[s], ((it - 4) >> 2) & 1);
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + s * chunk)), "l"(src + off), "r"(chunk), "r"(smem_u32(&bar[s])) : "memory");
        off += (size_t)chunk * 151;
        if (off + chunk > src_bytes) off = (off % (src_bytes - chunk)) & ~(size_t)1023;
      }
    }
    cycles_out[blockIdx.x] = clock64() - t0;
  }
}

static void probe_d(int num_sms) {
  printf("== probe D: L2->smem bulk-copy bandwidth, all SMs streaming ==\n");
  long long* d_cyc; uint8_t* d_src;
  CHECK(cudaMalloc(&d_cyc, num_sms * sizeof(long long)));
  int clock_khz = 0;
  CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  for (size_t src_mb : {32, 1024}) {
    const size_t bytes = src_mb << 20;
    CHECK(cudaMalloc(&d_src, bytes));
    CHECK(cudaMemset(d_src, 1, bytes));
    for (int chunk : {16384, 32768}) {
      const int smem = 4 * chunk + 1024;
      CHECK(cudaFuncSetAttribute(k_bulk_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int iters = 4096;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_bulk_bw<<<num_sms, 128, smem>>>(d_src, bytes, chunk, 64, d_cyc);   // warm
      cudaEventRecord(e0);
      k_bulk_bw<<<num_sms, 128, smem>>>(d_src, bytes, chunk, iters, d_cyc);
      cudaEventRecord(e1);
      CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cyc(num_sms);
      CHECK(cudaMemcpy(cyc.data(), d_cyc, num_sms * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0; for (long long c : cyc) mx = c > mx ? c : mx;
      const double total = (double)num_sms * iters * chunk;
      printf("  source %4zu MB, chunk %5d B: %.1f GB/s total, %.1f B/cycle/SM (max cycles %lld, %.3f ms)\n", src_mb, chunk,
             total / ms / 1e6, (double)iters * chunk / mx, mx, ms);
    }
    cudaFree(d_src);
  }
  cudaFree(d_cyc);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  const char* which = argc > 1 ? argv[1] : "abcd";
  for (const char* p = which; *p; ++p) {
    if (*p == 'a') probe_a();
    if (*p == 'b') probe_b();
    if (*p == 'c') probe_c(prop.multiProcessorCount);
    if (*p == 'd') probe_d(prop.multiProcessorCount);
    fflush(stdout);
  }
  return 0;
}
