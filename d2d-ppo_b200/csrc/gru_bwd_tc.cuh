// Fused back-propagation through the GRU window on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// For every row (agent g, time t, env b) the L steps of the training window are walked backwards with d(h) held in
// registers (reference: autograd through nn.GRU inside PPO.train_step, algorithms/d2d_ppo.py:198-216 with the
// windows of preprocess_input_for_rnn, :385-398):
//
//   dn  = dh (1 - z)(1 - n^2)      dz = dh (h_prev - n) z (1 - z)      dr = dn gh_n r (1 - r)
//   d(gh) = (dr, dz, dn r)         d(gi) = (dr, dz, dn)                dh_prev = dh z + d(gh) W_hh
//
// replaces, per step, gru_gate_bwd_kernel + the K = 3H data-gradient GEMM (6 + 5 launches per window) and their
// d(h) / d(gh) round trips through HBM.  Per row-step the kernel reads r, z, n, gh_n, h_prev (5H floats, written by
// the forward window kernel in store mode) and writes d(gh) in place of r, z, n (3H floats, consumed by the
// weight-gradient kernel); d(gi) of the observation the step looked at is accumulated with fp32 reductions
// (red.global.add: up to L windows share an observation).
//
// CTA = 16 warps on ONE 128-row tile (512 threads leave 128 registers each; a 17th warp would be charged as four):
// warp w serves TMEM lane quadrant w % 4 (row = 32 (w % 4) + lane) and the hidden units of quarter w / 4; lane 0 of
// warp 0 also issues the MMAs once every thread's operands are in place.  d(gh) is cut into three bf16 planes and written as the K-major A operand
// [128][3H]; W_hh^T (B operand [H][3H], three planes) stays resident; the six plane products with i + j <= 2 give
// fp32-level accuracy; D = d(gh) W_hh [128][H] lands in TMEM and is added to dh z by the thread that owns the row.
// Shared memory (H = 64): A 3 x 48 KB + W_hh^T 3 x 24 KB = 216 KB, hence one tile in flight; the next step's
// activations are requested before the thread waits for the MMA, so HBM stays busy while the tensor pipe works.
#pragma once
#include "gru_tc.cuh"

namespace d2d {

struct GruBwdTcArgs {
  View acts;   // [.. 4H ..] step 0; step s is acts_step floats further.  in: r, z, n, gh_n   out: dr, dz, dn r, (gh_n)
  View hs;     // [.. H ..]  h after step 0; step s is hs_step floats further
  View dh;     // [.. H ..]  in: d(loss) / d(h after the last step)
  View dgi;    // [.. 3H ..] accumulated at observation time t - (L - 1 - s); zeroed by the caller
  const float* w;
  long long w_agent_stride;
  int whh_off[D2D_MAX_AGENTS];
  long long acts_step, hs_step;
  int L, B, t0, t1;
};

namespace tcb {
constexpr int kThreads = 512;
constexpr int kGateThreads = 512;
template <int H>
struct Smem {
  static constexpr int kA = tc::kM * 3 * H;     // bf16 elements per A plane
  static constexpr int kW = H * 3 * H;          // bf16 elements per W_hh^T plane
  static constexpr size_t bytes = (size_t)(3 * kA + 3 * kW) * 2 + 64;
};
}  // namespace tcb

template <int H>
__global__ void __launch_bounds__(tcb::kThreads, 1) gru_bwd_tc_kernel(const GruBwdTcArgs a) {
  using namespace tc;
  using S = tcb::Smem<H>;
  static_assert(H == 32 || H == 64, "units per thread (H / 4) must be a multiple of 8");
  constexpr int UT = H / 4;      // hidden units per thread
  constexpr int K3 = 3 * H;      // reduction length of d(gh) W_hh
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [3 planes][128][3H]
  __nv_bfloat16* sw = sa + 3 * S::kA;                                   // [3 planes][H][3H]   (n = unit, k = gate row)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sw + 3 * S::kW);         // a_ready, d_ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y;
  const float* Whh = a.w + g * a.w_agent_stride + a.whh_off[g];        // [3H][H] row-major

  for (int i = tid; i < K3 * H; i += tcb::kThreads) {
    const int j = i / H, u = i % H;                                     // coalesced read of W_hh[j][u]
    __nv_bfloat16 p0, p1, p2;
    split3(Whh[i], p0, p1, p2);
    const int o = canon16(u, j, K3);
    sw[o] = p0, sw[S::kW + o] = p1, sw[2 * S::kW + o] = p2;
  }
  if (tid == 0) {
    mbar_init(a_ready, tcb::kGateThreads), mbar_init(d_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kCols = H < 32 ? 32 : H;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int L = a.L;
  const int n_t = a.t1 - a.t0;
  const int blocks = (a.B + kM - 1) / kM;
  const int n_tiles = n_t * blocks;     // tile p = (env block p / n_t, time t0 + p % n_t): the <= L tiles that touch
                                        // one d(gi) element are processed close together (L2-resident reductions)

  {
    // =================== gate warps ===================
    const int quarter = warp >> 2;
    const int row = ((warp & 3) << 5) + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    const int u0 = quarter * UT;
    const long long HB = (long long)H * a.B;
    uint32_t ph = 0, ph_a = 0;
    const uint32_t idesc = idesc_bf16(H);

    for (int p = blockIdx.x; p < n_tiles; p += gridDim.x) {
      const int t = a.t0 + p % n_t;
      const int b = (p / n_t) * kM + row;
      const bool ok = b < a.B;
      const int bb = ok ? b : 0;
      float dh[UT];
      float r[UT], z[UT], nn[UT], ghn[UT], hp[UT];
      {
        const float* dp = view_ptr(a.dh, g, t, a.B, bb) + (long long)u0 * a.B;
#pragma unroll
        for (int u = 0; u < UT; ++u) dh[u] = ok ? dp[(long long)u * a.B] : 0.f;
      }
      auto load_step = [&](int s) {
        const float* ap = view_ptr(a.acts, g, t, a.B, bb) + (long long)s * a.acts_step + (long long)u0 * a.B;
#pragma unroll
        for (int u = 0; u < UT; ++u) {
          const long long f = (long long)u * a.B;
          r[u] = ap[f], z[u] = ap[f + HB], nn[u] = ap[f + 2 * HB], ghn[u] = ap[f + 3 * HB];
        }
        if (s > 0) {
          const float* hq = view_ptr(a.hs, g, t, a.B, bb) + (long long)(s - 1) * a.hs_step + (long long)u0 * a.B;
#pragma unroll
          for (int u = 0; u < UT; ++u) hp[u] = hq[(long long)u * a.B];
        } else {
#pragma unroll
          for (int u = 0; u < UT; ++u) hp[u] = 0.f;
        }
      };
      load_step(L - 1);
      for (int s = L - 1; s >= 0; --s) {
        float* ap = view_ptr(a.acts, g, t, a.B, bb) + (long long)s * a.acts_step + (long long)u0 * a.B;
        float* gp = view_ptr(a.dgi, g, t - (L - 1 - s), a.B, bb) + (long long)u0 * a.B;
#pragma unroll
        for (int c = 0; c < UT / 8; ++c) {
          float dg[3][8];                      // dr, dz, dn r of 8 units
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int u = c * 8 + j;
            const float d = ok ? dh[u] : 0.f;
            const float dn = d * (1.0f - z[u]) * (1.0f - nn[u] * nn[u]);
            const float dz = d * (hp[u] - nn[u]) * z[u] * (1.0f - z[u]);
            const float dr = dn * ghn[u] * r[u] * (1.0f - r[u]);
            dg[0][j] = dr, dg[1][j] = dz, dg[2][j] = dn * r[u];
            dh[u] = d * z[u];                  // direct path; d(gh) W_hh is added once the MMA is done
            if (ok) {
              const long long f = (long long)u * a.B;
              ap[f] = dr, ap[f + HB] = dz, ap[f + 2 * HB] = dg[2][j];
              atomicAdd(gp + f, dr), atomicAdd(gp + f + HB, dz), atomicAdd(gp + f + 2 * HB, dn);
            }
          }
          if (s > 0) {
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
              uint32_t w0[4], w1[4], w2[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t a0, a1, a2, b0, b1, b2;
                split3_trunc(dg[gate][2 * j], a0, a1, a2);
                split3_trunc(dg[gate][2 * j + 1], b0, b1, b2);
                w0[j] = pack_hi(a0, b0), w1[j] = pack_hi(a1, b1), w2[j] = pack_hi(a2, b2);
              }
              __nv_bfloat16* dst = sa + canon16(row, gate * H + u0 + c * 8, K3);
              *reinterpret_cast<uint4*>(dst) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
              *reinterpret_cast<uint4*>(dst + S::kA) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
              *reinterpret_cast<uint4*>(dst + 2 * S::kA) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            }
          }
        }
        if (s == 0) break;                     // dh of the zero initial state is not needed
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(a_ready);
        if (tid == 0) {                        // MMA issue: six plane products x 3H / 16 reduction steps
          mbar_wait(a_ready, ph_a);
          ph_a ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          bool first = true;
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (i + j > 2) continue;
              const uint32_t aa = smem_u32(sa + i * S::kA);
              const uint32_t wb = smem_u32(sw + j * S::kW);
#pragma unroll
              for (int k16 = 0; k16 < K3 / 16; ++k16) {
                mma_bf16(tmem, desc16(aa + k16 * 256, K3), desc16(wb + k16 * 256, K3), idesc, !first);
                first = false;
              }
            }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                       ::"r"(smem_u32(d_ready))
                       : "memory");
        }
        __syncwarp();
        load_step(s - 1);                      // in flight while the tensor pipe works
        mbar_wait(d_ready, ph);
        ph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c = 0; c < UT / 8; ++c) {
          float v[8];
          tmem_ld8(tmem + lane_addr + (uint32_t)(u0 + c * 8), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) dh[c * 8 + j] += v[j];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols));
}

}  // namespace d2d
