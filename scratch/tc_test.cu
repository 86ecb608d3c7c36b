// Standalone check of the tcgen05 building blocks used by the learner kernels:
// D[128 x N] = A[128 x K] * B[N x K]^T with 3xTF32 split accumulation, A/B in shared memory (K-major, no swizzle),
// accumulator in TMEM, read back with tcgen05.ld.  Compares against a float64 CPU result.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major / no-swizzle tile [rows][kc]: 8 x 16-byte core matrices, K-adjacent cores contiguous
__device__ __host__ __forceinline__ int canon_off(int r, int k, int kc) {   // in floats
  return ((r >> 3) * (kc >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int kc) {
  const uint64_t start = (saddr & 0x3FFFFu) >> 4;
  const uint64_t lbo = 128 >> 4;                 // next core matrix along K
  const uint64_t sbo = ((kc >> 2) * 128) >> 4;   // next 8-row group
  return start | (lbo << 16) | (sbo << 32) | (1ull << 46);
}

__device__ __forceinline__ float tf32_hi(float x) {   // round to nearest tf32
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(128) tc_gemm(const float* __restrict__ A, const float* __restrict__ B, float* D, int N,
                                               int K, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem);
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;
  float* b_lo = b_hi + N * K;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + N * K);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const float v = A[i], h = tf32_hi(v);
    a_hi[canon_off(r, k, K)] = h;
    a_lo[canon_off(r, k, K)] = tf32_hi(v - h);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const float v = B[i], h = tf32_hi(v);
    b_hi[canon_off(r, k, K)] = h;
    b_lo[canon_off(r, k, K)] = tf32_hi(v - h);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t accum = 0;
    for (int p = 0; p < passes; ++p) {
      const float* ap = (p == 1) ? a_lo : a_hi;
      const float* bp = (p == 2) ? b_lo : b_hi;
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint64_t ad = make_desc(smem_u32(ap) + k8 * 256, K);
        const uint64_t bd = make_desc(smem_u32(bp) + k8 * 256, K);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum)
            : "memory");
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                 : "memory");
  }
  // wait for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(mbar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32); thread = row
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ---- bf16 x 3 planes, kind::f16: products (i, j) with i + j <= max_order ----
__device__ __host__ __forceinline__ int canon_off16(int r, int k, int kc) {   // in bf16 elements
  return ((r >> 3) * (kc >> 3) + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7);
}
__device__ __forceinline__ uint64_t make_desc16(uint32_t saddr, int kc) {
  const uint64_t start = (saddr & 0x3FFFFu) >> 4;
  const uint64_t lbo = 128 >> 4, sbo = ((kc >> 3) * 128) >> 4;
  return start | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
__device__ __forceinline__ void split3(float v, __nv_bfloat16* p) {
  p[0] = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(p[0]);
  p[1] = __float2bfloat16_rn(r);
  r -= __bfloat162float(p[1]);
  p[2] = __float2bfloat16_rn(r);
}

__global__ void __launch_bounds__(128) tc_gemm_bf16(const float* __restrict__ A, const float* __restrict__ B, float* D,
                                                    int N, int K, int max_order) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* a_pl = reinterpret_cast<__nv_bfloat16*>(smem);   // [3][M*K]
  __nv_bfloat16* b_pl = a_pl + 3 * M * K;                         // [3][N*K]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_pl + 3 * N * K);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += 128) {
    __nv_bfloat16 p[3];
    split3(A[i], p);
    for (int q = 0; q < 3; ++q) a_pl[q * M * K + canon_off16(i / K, i % K, K)] = p[q];
  }
  for (int i = tid; i < N * K; i += 128) {
    __nv_bfloat16 p[3];
    split3(B[i], p);
    for (int q = 0; q < 3; ++q) b_pl[q * N * K + canon_off16(i / K, i % K, K)] = p[q];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t accum = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        if (i + j > max_order) continue;
        for (int k16 = 0; k16 < K / 16; ++k16) {
          const uint64_t ad = make_desc16(smem_u32(a_pl + i * M * K) + k16 * 256, K);
          const uint64_t bd = make_desc16(smem_u32(b_pl + j * N * K) + k16 * 256, K);
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum)
              : "memory");
          accum = 1;
        }
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                 : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(mbar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 192, K = argc > 2 ? atoi(argv[2]) : 64;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4), cudaMalloc(&dB, B.size() * 4), cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4 + 64;
  cudaFuncSetAttribute(tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int passes = 1; passes <= 3; passes += 2) {
    cudaMemset(dD, 0, D.size() * 4);
    tc_gemm<<<1, 128, smem>>>(dA, dB, dD, N, K, passes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
        maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
        maxref = fmax(maxref, fabs(s));
      }
    printf("N=%d K=%d passes=%d: max|err| = %.3e, max|ref| = %.3f, rel = %.3e\n", N, K, passes, maxerr, maxref,
           maxerr / maxref);
  }
  const size_t smem16 = (size_t)(3 * M * K + 3 * N * K) * 2 + 64;
  cudaFuncSetAttribute(tc_gemm_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);
  for (int order = 0; order <= 2; ++order) {
    cudaMemset(dD, 0, D.size() * 4);
    tc_gemm_bf16<<<1, 128, smem16>>>(dA, dB, dD, N, K, order);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error (bf16): %s\n", cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
        maxerr = fmax(maxerr, fabs(s - D[m * N + n]));
        maxref = fmax(maxref, fabs(s));
      }
    printf("bf16 planes, products with i+j <= %d: max|err| = %.3e, rel = %.3e\n", order, maxerr, maxerr / maxref);
  }
  return 0;
}
