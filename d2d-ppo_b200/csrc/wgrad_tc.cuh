// Weight gradients on the tensor cores (tcgen05 + TMEM), sm_100a:
//
//   dW[o][k] = sum_rows dy[o][row] * x[k][row],   db[o] = sum_rows dy[o][row]      (per agent, partial per CTA)
//
// This is the one learner GEMM whose operands are K-major straight out of the env-minor layout: the reduction
// index is the row (time, env), and rows are the fastest dimension of every activation matrix.  The reduction is
// long (all rows of the batch), so it is a classic pipelined GEMM: 8 producer warps stream 32-row chunks of dy and x,
// cut each fp32 value into three bf16 planes (v = p0 + p1 + p2) and store them in the canonical K-major UMMA layout;
// one thread issues, per chunk, the six plane products with i + j <= 2 (fp32-level accuracy) for up to two M = 128
// blocks of outputs; the fp32 accumulators stay in TMEM for the whole kernel and are read once at the end.
// The bias gradient is the extra B row of ones.  Split-K over CTAs: every CTA writes a partial that
// wreduce_kernel sums in a fixed order (deterministic).
#pragma once
#include "gru_tc.cuh"

namespace d2d {
namespace tcw {

constexpr int kRows = 32;        // rows (reduction elements) per chunk = two k16 MMA steps
constexpr int kThreads = 288;    // 8 producer warps + 1 MMA warp
constexpr int kProducers = 256;

// A planes hold 192 real feature rows; a second M = 128 block (features 128..255) reads up to 64 rows past them,
// which lands in the B planes of the same stage: finite garbage feeding accumulator rows that are never read.
constexpr int kARows = 192;
constexpr int kAPlane = kARows * kRows;    // bf16 elements

inline __host__ __device__ int b_rows(int K) { return (K + 1 + 15) / 16 * 16; }   // + ones row, N % 16 == 0
inline __host__ __device__ size_t stage_elems(int K) { return (size_t)3 * kAPlane + 3 * b_rows(K) * kRows; }
// + 8 KB slack: with a narrow B tile (K <= 15) the second output block's reads run past the end of stage 1
inline __host__ __device__ size_t smem_bytes(int K) { return 2 * stage_elems(K) * 2 + 64 + 8192; }

__device__ __forceinline__ uint32_t idesc_bf16_m128(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace tcw

// requires B % 8 == 0, out_dim <= 192, in_dim <= 128
__global__ void __launch_bounds__(tcw::kThreads, 2) wgrad_tc_kernel(const WgradArgs a) {
  using tc::canon16; using tc::desc16; using tc::mbar_arrive; using tc::mbar_init; using tc::mbar_wait;
  using tc::mma_bf16; using tc::pack_hi; using tc::smem_u32; using tc::split3_trunc; using tc::tmem_ld8;
  using namespace tcw;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int g = blockIdx.y, strip = blockIdx.x, n_strips = gridDim.x;
  const int K = a.in_dim[g], O = a.out_dim;
  const int NB = b_rows(K);                    // B tile rows = MMA N
  const size_t st_elems = stage_elems(K);
  __nv_bfloat16* stage_base = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + 2 * st_elems);   // full[2], empty[2], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  uint64_t *full = bars, *empty = bars + 2, *done = bars + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blocks = O > 128 ? 2 : 1;

  const int chunks_per_t = (a.B + kRows - 1) / kRows;
  const long long n_chunks = (long long)(a.t1 - a.t0) * chunks_per_t;
  const int my_chunks = strip < n_chunks ? (int)((n_chunks - 1 - strip) / n_strips) + 1 : 0;

  // zero both stages once: feature rows >= O / >= K + 1 stay zero for the whole kernel
  for (size_t i = tid; i < 2 * st_elems / 8; i += kThreads) reinterpret_cast<uint4*>(stage_base)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&full[0], kProducers), mbar_init(&full[1], kProducers);
    mbar_init(&empty[0], 1), mbar_init(&empty[1], 1), mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // accumulator columns: n_blocks x NB, rounded up to a power of two >= 32 (two CTAs per SM share the 512 columns)
  uint32_t n_cols = 32;
  while (n_cols < (uint32_t)(n_blocks * NB)) n_cols <<= 1;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(n_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =================== MMA issuer ===================
    if (lane == 0 && my_chunks > 0) {
      const uint32_t idesc = idesc_bf16_m128(NB);
      const uint32_t sbo = (kRows >> 3) * 128;   // bytes between 8-row groups of a [.][32] tile
      const int bp = a.x_planes > 0 ? a.x_planes : 3;   // x exact in bf16: planes 1, 2 of B are zero, their products too
      bool first = true;
      for (int i = 0; i < my_chunks; ++i) {
        const int st = i & 1;
        mbar_wait(&full[st], (uint32_t)((i >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const __nv_bfloat16* sa = stage_base + st * st_elems;
        const __nv_bfloat16* sb = sa + 3 * kAPlane;
#pragma unroll
        for (int pi = 0; pi < 3; ++pi)
#pragma unroll
          for (int pj = 0; pj < 3; ++pj) {
            if (pi + pj > 2 || pj >= bp) continue;
            const uint32_t aa = smem_u32(sa + pi * kAPlane);
            const uint32_t bb = smem_u32(sb + pj * NB * kRows);
#pragma unroll
            for (int k16 = 0; k16 < kRows / 16; ++k16) {
              const uint64_t bd = desc16(bb + k16 * 256, kRows);
              for (int blk = 0; blk < n_blocks; ++blk)
                mma_bf16(tmem + (uint32_t)(blk * NB), desc16(aa + blk * 16 * sbo + k16 * 256, kRows), bd, idesc, !first);
              first = false;
            }
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(smem_u32(&empty[st]))
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done))
                   : "memory");
    }
  } else {
    // =================== producers: (feature, 8-row group) items -> three bf16 planes ===================
    const int n_feat = O + K + 1;              // dy features, x features, the ones row
    const int n_items = n_feat * (kRows / 8);
    for (int i = 0; i < my_chunks; ++i) {
      const int st = i & 1;
      if (i >= 2) mbar_wait(&empty[st], (uint32_t)(((i >> 1) - 1) & 1));
      const long long c = strip + (long long)i * n_strips;
      const int t = a.t0 + (int)(c / chunks_per_t);
      const int b0 = (int)(c % chunks_per_t) * kRows;
      __nv_bfloat16* sa = stage_base + st * st_elems;
      __nv_bfloat16* sb = sa + 3 * kAPlane;
      // every thread owns up to kItems (feature, 8-row group) items of the chunk; all of their loads are issued before
      // the first use, so a producer keeps kItems x 32 bytes in flight instead of one item's
      constexpr int kItems = 4;                // typical chunk: (192 + 64 + 1) features x 4 groups / 256 threads = 4.02
      for (int base = 0; base < n_items; base += kItems * kProducers) {
      float4 lo[kItems], hi[kItems];
#pragma unroll
      for (int q = 0; q < kItems; ++q) {
        const int it = base + tid + q * kProducers;
        lo[q] = hi[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < n_items) {
          const int f = it / (kRows / 8), r8 = (it % (kRows / 8)) * 8;
          const bool ok = b0 + r8 < a.B;        // B % 8 == 0: a group of 8 rows is all valid or all past the end
          const float* src = nullptr;
          if (ok && f < O) src = view_ptr(a.dy, g, t, a.B, b0 + r8) + (long long)f * a.B;
          else if (ok && f - O < K) src = view_ptr(a.x, g, t, a.B, b0 + r8) + (long long)(f - O) * a.B;
          if (src) {
            lo[q] = __ldg(reinterpret_cast<const float4*>(src));
            hi[q] = __ldg(reinterpret_cast<const float4*>(src) + 1);
          } else if (ok && f - O == K) {
            lo[q] = hi[q] = make_float4(1.f, 1.f, 1.f, 1.f);   // the ones row: db = sum of dy
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kItems; ++q) {
        const int it = base + tid + q * kProducers;
        if (it < n_items) {
          const int f = it / (kRows / 8), r8 = (it % (kRows / 8)) * 8;
          __nv_bfloat16* dst;
          int plane_stride;
          if (f < O) dst = sa + canon16(f, r8, kRows), plane_stride = kAPlane;
          else dst = sb + canon16(f - O, r8, kRows), plane_stride = NB * kRows;
          const float v[8] = {lo[q].x, lo[q].y, lo[q].z, lo[q].w, hi[q].x, hi[q].y, hi[q].z, hi[q].w};
          uint32_t q0[8], q1[8], q2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split3_trunc(v[j], q0[j], q1[j], q2[j]);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_hi(q0[0], q0[1]), pack_hi(q0[2], q0[3]), pack_hi(q0[4], q0[5]), pack_hi(q0[6], q0[7]));
          if (f >= O && a.x_planes == 1) continue;     // exact x: the lower planes stay at the zeros of the initial fill
          *reinterpret_cast<uint4*>(dst + plane_stride) =
              make_uint4(pack_hi(q1[0], q1[1]), pack_hi(q1[2], q1[3]), pack_hi(q1[4], q1[5]), pack_hi(q1[6], q1[7]));
          *reinterpret_cast<uint4*>(dst + 2 * plane_stride) =
              make_uint4(pack_hi(q2[0], q2[1]), pack_hi(q2[2], q2[3]), pack_hi(q2[4], q2[5]), pack_hi(q2[6], q2[7]));
        }
      }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&full[st]);
    }
    // =================== epilogue: warps 0-3 read the accumulators (one TMEM lane quadrant each) ===================
    if (warp < 4) {
      float* out = a.partial + ((long long)g * n_strips + strip) * a.part_stride;
      if (my_chunks > 0) {
        mbar_wait(done, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int o = blk * 128 + warp * 32 + lane;
        const uint32_t d = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * NB);
        for (int c0 = 0; c0 < NB; c0 += 8) {
          float v[8];
          if (my_chunks > 0) {
            tmem_ld8(d + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
          if (o < O) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = c0 + j;
              if (k < K) out[(long long)o * K + k] = v[j];
              else if (k == K && a.with_bias) out[(long long)O * K + o] = v[j];
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(n_cols));
}

}  // namespace d2d
