// Grouped dense layer on the tensor cores (tcgen05 + TMEM), sm_100a: the DenseArgs contract of dense_kernel,
//   y[out][row] = epilogue( sum_in x[in][row] * M[in][out] + bias[out] ),   rows tiled 128 at a time,
// used for the input projection of training windows (K = I, N = 3H), the data gradients (K = 3H, N = H), and the
// network heads (K = H, N = H / O).
//
// Activations are row-contiguous per feature, i.e. M-major for a row-tile GEMM; every gate thread owns one row,
// gathers 8 consecutive features (8 coalesced loads across the warp), cuts them into three bf16 planes and writes
// one 16-byte vector per plane into the canonical K-major UMMA layout: the transposition costs nothing extra.
// Weights (three bf16 planes) stay resident in shared memory; accumulators live in TMEM; plane products with
// i + j <= 2 give fp32-level accuracy (inputs flagged exact-in-bf16 need one A plane only).
// Two 128-row slots per CTA overlap staging / epilogue of one tile with the MMAs of the other, as in gru_tc.cuh.
#pragma once
#include "gru_tc.cuh"

namespace d2d {
namespace tcd {

constexpr int kThreads = 288;   // warps 0-3: slot 0, warps 4-7: slot 1 (thread = row), warp 8: MMA issuer
constexpr int kKc = 64;         // input features staged per chunk

inline __host__ __device__ int pad16(int v) { return (v + 15) / 16 * 16; }
inline __host__ __device__ size_t smem_bytes(int in_dim, int out_dim) {
  const int kp = pad16(in_dim > 0 ? in_dim : 1), np = pad16(out_dim);
  const int kc = kp < kKc ? kp : kKc;
  const int kpad = (kp + kc - 1) / kc * kc;   // whole chunks
  return (size_t)(3 * np * kpad + 2 * 3 * 128 * kc) * 2 + (size_t)np * 4 + 64;
}

}  // namespace tcd

// requires out_dim <= 192 (two slots x N columns <= 512 TMEM columns after rounding), in_dim <= 192
__global__ void __launch_bounds__(tcd::kThreads, 1) dense_tc_kernel(const DenseArgs a, const int x_exact) {
  using tc::canon16; using tc::desc16; using tc::mbar_arrive; using tc::mbar_init; using tc::mbar_wait;
  using tc::mma_bf16; using tc::pack_hi; using tc::smem_u32; using tc::split3_trunc; using tc::tmem_ld8;
  using namespace tcd;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int g = blockIdx.y;
  const int in_dim = a.in_dim[g], out_dim = a.out_dim;
  const int KP = pad16(in_dim > 0 ? in_dim : 1), NP = pad16(out_dim);
  const int KC = KP < kKc ? KP : kKc;              // features per chunk (multiple of 16)
  const int n_kc = (KP + KC - 1) / KC;
  const int KPAD = n_kc * KC;
  __nv_bfloat16* wpl = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [3 planes][n_kc chunks][NP][KC]
  __nv_bfloat16* apl = wpl + 3 * NP * KPAD;                          // [2 slots][3 planes][128][KC]
  float* bias = reinterpret_cast<float*>(apl + 2 * 3 * 128 * KC);    // [NP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias + NP);           // a_ready[2], d_ready[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint64_t *a_ready = bars, *d_ready = bars + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_planes_a = x_exact ? 1 : 3;

  // ---- weights -> bf16 planes, chunked along K so that every chunk is one canonical [NP][KC] tile ----
  const float* W = a.w + g * a.w_agent_stride + a.w_off[g];
  const int ld = a.w_ld[g];
  const int plane = NP * KPAD;
  for (int i = tid; i < NP * KPAD; i += kThreads) {
    const int n = i / KPAD, k = i % KPAD;
    float v = 0.f;
    if (n < out_dim && k < in_dim) v = a.trans ? W[(long long)k * ld + n] : W[(long long)n * ld + k];
    uint32_t t0, t1, t2;
    split3_trunc(v, t0, t1, t2);
    const int o = (k / KC) * NP * KC + canon16(n, k % KC, KC);
    wpl[o] = __ushort_as_bfloat16((unsigned short)(t0 >> 16));
    wpl[plane + o] = __ushort_as_bfloat16((unsigned short)(t1 >> 16));
    wpl[2 * plane + o] = __ushort_as_bfloat16((unsigned short)(t2 >> 16));
  }
  for (int o = tid; o < NP; o += kThreads)
    bias[o] = (o < out_dim && a.b_off[g] >= 0) ? a.w[g * a.w_agent_stride + a.b_off[g] + o] : 0.f;
  if (tid == 0) {
    mbar_init(&a_ready[0], 128), mbar_init(&a_ready[1], 128);
    mbar_init(&d_ready[0], 1), mbar_init(&d_ready[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t n_cols = 32;
  while (n_cols < (uint32_t)(2 * NP)) n_cols <<= 1;
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

  const int pairs_per_t = (a.B + 255) / 256;
  const int n_pairs = (a.t1 - a.t0) * pairs_per_t;

  if (warp == 8) {
    // =================== MMA issuer ===================
    if (lane == 0) {
      uint32_t ph[2] = {0, 0};
      const uint32_t idesc = tc::idesc_bf16(NP);
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int kc = 0; kc < n_kc; ++kc) {
          for (int slot = 0; slot < 2; ++slot) {
            mbar_wait(&a_ready[slot], ph[slot]);
            ph[slot] ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem + (uint32_t)(slot * NP);
            bool first = kc == 0;
            for (int i = 0; i < n_planes_a; ++i)
              for (int j = 0; j < 3; ++j) {
                if (i + j > 2) continue;
                const uint32_t aa = smem_u32(apl + (slot * 3 + i) * 128 * KC);
                const uint32_t bb = smem_u32(wpl + j * plane + kc * NP * KC);
                for (int k16 = 0; k16 < KC / 16; ++k16) {
                  mma_bf16(d, desc16(aa + k16 * 256, KC), desc16(bb + k16 * 256, KC), idesc, !first);
                  first = false;
                }
              }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&d_ready[slot]))
                         : "memory");
          }
        }
      }
    }
  } else {
    // =================== row threads: stage A chunks, epilogue ===================
    const int slot = warp >> 2;
    const int row = ((warp & 3) << 5) + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    __nv_bfloat16* my_a = apl + slot * 3 * 128 * KC;
    uint32_t ph = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const int t = a.t0 + p / pairs_per_t;
      const int b = (p % pairs_per_t) * 256 + slot * 128 + row;
      const bool ok = b < a.B;
      const float* xp = view_ptr(a.x, g, t, a.B, ok ? b : 0);
      for (int kc = 0; kc < n_kc; ++kc) {
        // the previous chunk's MMAs (which read this slot's A tile) must be done before it is overwritten
        if (kc > 0) {
          mbar_wait(&d_ready[slot], ph);
          ph ^= 1u;
        }
        // all loads of the chunk are issued before the first use (up to 64 independent coalesced loads in flight)
        float xv[kKc];
#pragma unroll
        for (int kk = 0; kk < kKc; ++kk) {
          const int k = kc * KC + kk;
          xv[kk] = (kk < KC && ok && k < in_dim) ? xp[(long long)k * a.B] : 0.f;
        }
#pragma unroll
        for (int k8 = 0; k8 < kKc; k8 += 8) {
          if (k8 < KC) {
            uint32_t q0[8], q1[8], q2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split3_trunc(xv[k8 + j], q0[j], q1[j], q2[j]);
            const int o = canon16(row, k8, KC);
            *reinterpret_cast<uint4*>(my_a + o) =
                make_uint4(pack_hi(q0[0], q0[1]), pack_hi(q0[2], q0[3]), pack_hi(q0[4], q0[5]), pack_hi(q0[6], q0[7]));
            if (!x_exact) {
              *reinterpret_cast<uint4*>(my_a + 128 * KC + o) =
                  make_uint4(pack_hi(q1[0], q1[1]), pack_hi(q1[2], q1[3]), pack_hi(q1[4], q1[5]), pack_hi(q1[6], q1[7]));
              *reinterpret_cast<uint4*>(my_a + 2 * 128 * KC + o) =
                  make_uint4(pack_hi(q2[0], q2[1]), pack_hi(q2[2], q2[3]), pack_hi(q2[4], q2[5]), pack_hi(q2[6], q2[7]));
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_ready[slot]);
      }
      float* yp = view_ptr(a.y, g, t, a.B, ok ? b : 0);
      const float* ap = a.epilogue == kEpiReluBwd ? view_ptr(a.aux, g, t, a.B, ok ? b : 0) : nullptr;
      // the epilogue's second operand (ReLU mask / accumulation target) is requested before the wait for the MMAs
      constexpr int kPre = 64;
      const bool pre_ok = NP <= kPre && (a.epilogue == kEpiReluBwd || a.epilogue == kEpiAccum);
      float pre[kPre];
      if (pre_ok) {
        const float* src = a.epilogue == kEpiReluBwd ? ap : yp;
#pragma unroll
        for (int o = 0; o < kPre; ++o) pre[o] = (ok && o < out_dim) ? src[(long long)o * a.B] : 0.f;
      }
      mbar_wait(&d_ready[slot], ph);
      ph ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem + (uint32_t)(slot * NP) + lane_addr;
      if (pre_ok) {
#pragma unroll
        for (int c0 = 0; c0 < kPre; c0 += 8) {
          if (c0 < NP) {
            float v[8];
            tmem_ld8(d + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int o = c0 + j;
                if (o < out_dim) {
                  float r = v[j] + bias[o];
                  if (a.epilogue == kEpiAccum) r += pre[o];
                  else r = pre[o] > 0.f ? r : 0.f;
                  yp[(long long)o * a.B] = r;
                }
              }
            }
          }
        }
      } else
      for (int c0 = 0; c0 < NP; c0 += 8) {
        float v[8];
        tmem_ld8(d + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int o = c0 + j;
            if (o < out_dim) {
              float r = v[j] + bias[o];
              float* q = yp + (long long)o * a.B;
              if (a.epilogue == kEpiRelu) r = fmaxf(r, 0.f);
              else if (a.epilogue == kEpiAccum) r += *q;
              else if (a.epilogue == kEpiReluBwd) r = ap[(long long)o * a.B] > 0.f ? r : 0.f;
              *q = r;
            }
          }
        }
      }
      // our TMEM reads are ordered before the next tile's first MMA by the fence in front of the next arrive
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(n_cols));
}

}  // namespace d2d
