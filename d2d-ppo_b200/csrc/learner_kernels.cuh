// Device kernels of the PPO / IPPO learner: per-agent networks evaluated for all N agents in one launch.
//
// Data layout ("env-minor", as everywhere in this library): an activation matrix is stored as
//   A[time block t][feature row f][env b]   with b fastest,
// so a warp of 32 consecutive envs reads / writes 32 consecutive floats for any (t, f).  Agent g's features
// start at row f_off[g] of a time block.  Every thread owns ONE row (t, b) of ONE agent: all per-row work
// (dot products against weights broadcast from shared memory, gate maths, distribution maths) needs no
// cross-thread traffic; only the weight gradients reduce over rows (wgrad_kernel).
//
// The B*N-row "MLP" of the north_star is N grouped GEMMs with K in {30, 64, 119} and fp32 parity at 1e-5, so this
// first implementation runs them as FP32 FMA on the CUDA cores (see DESIGN.md section 6 for the tcgen05 plan).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace d2d {

constexpr int kRowsPerBlock = 128;

struct View {
  float* p;
  long long t_stride;  // floats between consecutive time blocks
  int t_off;           // added to the time index (window shifts / halos)
  int f_off[D2D_MAX_AGENTS];
};

__device__ __forceinline__ float* view_ptr(const View& v, int g, int t, int B, int b) {
  return v.p + (long long)(t + v.t_off) * v.t_stride + (long long)v.f_off[g] * B + b;
}

// ------------------------------------------------------------------------------------------------
// dense: y[out][row] = epilogue( sum_in x[in][row] * M[in][out] + bias[out] )
//   trans = 0: M[in][out] = W[out][in]  (forward,   W stored [out_dim][in_dim], row length w_ld)
//   trans = 1: M[in][out] = W[in][out]  (backward-data through W stored [in_dim][out_dim], row length w_ld)
// ------------------------------------------------------------------------------------------------
enum { kEpiNone = 0, kEpiRelu = 1, kEpiAccum = 2, kEpiReluBwd = 3 };

struct DenseArgs {
  View x, y, aux;
  const float* w;
  long long w_agent_stride;
  int w_off[D2D_MAX_AGENTS];
  int b_off[D2D_MAX_AGENTS];  // < 0: no bias
  int in_dim[D2D_MAX_AGENTS];
  int w_ld[D2D_MAX_AGENTS];
  int out_dim;
  int trans;
  int epilogue;
  int B, t0, t1;
};

template <int OC>  // outputs accumulated per pass (registers per thread)
__global__ void __launch_bounds__(kRowsPerBlock) dense_kernel(const DenseArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y;
  const int in_dim = a.in_dim[g], out_dim = a.out_dim;
  const int out_pad = (out_dim + OC - 1) / OC * OC;
  float* Ms = sm;                      // [in_dim][out_pad]
  float* bs = sm + in_dim * out_pad;   // [out_pad]
  const float* W = a.w + g * a.w_agent_stride + a.w_off[g];
  const int ld = a.w_ld[g];
  for (int i = threadIdx.x; i < in_dim * out_pad; i += blockDim.x) {
    const int in = i / out_pad, o = i % out_pad;
    float v = 0.f;
    if (o < out_dim) v = a.trans ? W[(long long)in * ld + o] : W[(long long)o * ld + in];
    Ms[i] = v;
  }
  for (int o = threadIdx.x; o < out_pad; o += blockDim.x)
    bs[o] = (o < out_dim && a.b_off[g] >= 0) ? a.w[g * a.w_agent_stride + a.b_off[g] + o] : 0.f;
  __syncthreads();

  const int tiles_per_t = (a.B + kRowsPerBlock - 1) / kRowsPerBlock;
  const int n_tiles = (a.t1 - a.t0) * tiles_per_t;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int t = a.t0 + tile / tiles_per_t;
    const int b = (tile % tiles_per_t) * kRowsPerBlock + threadIdx.x;
    if (b >= a.B) continue;
    const float* xp = view_ptr(a.x, g, t, a.B, b);
    float* yp = view_ptr(a.y, g, t, a.B, b);
    const float* ap = a.epilogue == kEpiReluBwd ? view_ptr(a.aux, g, t, a.B, b) : nullptr;
    for (int o0 = 0; o0 < out_dim; o0 += OC) {
      float acc[OC];
#pragma unroll
      for (int j = 0; j < OC; ++j) acc[j] = bs[o0 + j];
      const float* mrow = Ms + o0;
#pragma unroll 4
      for (int in = 0; in < in_dim; ++in) {
        const float xv = xp[(long long)in * a.B];
        const float4* m4 = reinterpret_cast<const float4*>(mrow + in * out_pad);
#pragma unroll
        for (int j = 0; j < OC / 4; ++j) {
          const float4 m = m4[j];
          acc[4 * j + 0] = fmaf(xv, m.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(xv, m.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xv, m.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(xv, m.w, acc[4 * j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < OC; ++j) {
        const int o = o0 + j;
        if (o < out_dim) {
          float v = acc[j];
          float* q = yp + (long long)o * a.B;
          if (a.epilogue == kEpiRelu) v = fmaxf(v, 0.f);
          else if (a.epilogue == kEpiAccum) v += *q;
          else if (a.epilogue == kEpiReluBwd) v = ap[(long long)o * a.B] > 0.f ? v : 0.f;
          *q = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dense, register-tiled: the same contract as dense_kernel for B % 4 == 0.
// Block = 8 warps; every warp works on the SAME 32 * RPT rows (lane l owns the row quads [128 q + 4 l, +4), q < RPT/4:
// a 16-byte lane stride keeps the LDS.128 of the row tile conflict-free) and
// warp w owns outputs [w * OPT, w * OPT + OPT).  Per input k a thread issues RPT/4 LDS.128 for its rows and OPT/4
// broadcast LDS.128 for its weights against RPT * OPT FMAs (1 : 14 at <4, 24>, vs 1 : 4 for one row per thread).
// Input rows are streamed global -> shared with cp.async in chunks of 32 inputs, double buffered across
// (row tile, input chunk), so the global latency hides behind the FMAs of the previous chunk.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kDenseKC = 32;

template <int RPT, int OPT>
__global__ void __launch_bounds__(256) dense_tile_kernel(const DenseArgs a) {
  constexpr int ROWS = 32 * RPT, KC = kDenseKC, OUT_PAD = 8 * OPT;
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y;
  const int in_dim = a.in_dim[g], out_dim = a.out_dim;
  float* Ms = sm;                               // [in_dim][OUT_PAD]
  float* bs = Ms + (size_t)max(in_dim, 1) * OUT_PAD;   // [OUT_PAD]
  float* Xs = bs + OUT_PAD;                     // [2][KC][ROWS]
  const float* W = a.w + g * a.w_agent_stride + a.w_off[g];
  const int ld = a.w_ld[g];
  for (int i = threadIdx.x; i < in_dim * OUT_PAD; i += 256) {
    const int in = i / OUT_PAD, o = i % OUT_PAD;
    float v = 0.f;
    if (o < out_dim) v = a.trans ? W[(long long)in * ld + o] : W[(long long)o * ld + in];
    Ms[i] = v;
  }
  for (int o = threadIdx.x; o < OUT_PAD; o += 256)
    bs[o] = (o < out_dim && a.b_off[g] >= 0) ? a.w[g * a.w_agent_stride + a.b_off[g] + o] : 0.f;

  const int lane = threadIdx.x & 31, og = threadIdx.x >> 5;
  const int tiles_per_t = (a.B + ROWS - 1) / ROWS;
  const int n_tiles = (a.t1 - a.t0) * tiles_per_t;
  const int nkc = max(1, (in_dim + KC - 1) / KC);
  const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int total = my_tiles * nkc;

  auto issue = [&](int c) {   // chunk c -> buffer c & 1
    const int tile = blockIdx.x + (c / nkc) * gridDim.x;
    const int kc = c % nkc;
    const int t = a.t0 + tile / tiles_per_t, b0 = (tile % tiles_per_t) * ROWS;
    float* dst = Xs + (size_t)(c & 1) * KC * ROWS;
    const float* src = view_ptr(a.x, g, t, a.B, b0);
    for (int i = threadIdx.x; i < KC * (ROWS / 4); i += 256) {
      const int kk = i / (ROWS / 4), r4 = (i % (ROWS / 4)) * 4;
      const int k = kc * KC + kk;
      float* d = dst + kk * ROWS + r4;
      if (k < in_dim && b0 + r4 < a.B) cp_async16(d, src + (long long)k * a.B + r4);
      else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  if (total > 0) issue(0);
  float acc[RPT][OPT];
  for (int c = 0; c < total; ++c) {
    if (c + 1 < total) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // chunk c (and, the first time, the staged weights) visible to every warp
    const int kc = c % nkc;
    if (kc == 0) {
#pragma unroll
      for (int j = 0; j < OPT; ++j) {
        const float bv = bs[og * OPT + j];
#pragma unroll
        for (int r = 0; r < RPT; ++r) acc[r][j] = bv;
      }
    }
    const float* xs = Xs + (size_t)(c & 1) * KC * ROWS + lane * 4;   // quad q of this lane: rows 128 q + 4 lane
    const int kmax = min(KC, in_dim - kc * KC);
    const float* mrow = Ms + (size_t)kc * KC * OUT_PAD + og * OPT;
#pragma unroll 2
    for (int kk = 0; kk < kmax; ++kk) {
      float xv[RPT];
#pragma unroll
      for (int q = 0; q < RPT / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(xs + kk * ROWS + 128 * q);
        xv[4 * q] = v.x, xv[4 * q + 1] = v.y, xv[4 * q + 2] = v.z, xv[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int q = 0; q < OPT / 4; ++q) {
        const float4 m = *reinterpret_cast<const float4*>(mrow + kk * OUT_PAD + 4 * q);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          acc[r][4 * q + 0] = fmaf(xv[r], m.x, acc[r][4 * q + 0]);
          acc[r][4 * q + 1] = fmaf(xv[r], m.y, acc[r][4 * q + 1]);
          acc[r][4 * q + 2] = fmaf(xv[r], m.z, acc[r][4 * q + 2]);
          acc[r][4 * q + 3] = fmaf(xv[r], m.w, acc[r][4 * q + 3]);
        }
      }
    }
    if (kc == nkc - 1) {
      const int tile = blockIdx.x + (c / nkc) * gridDim.x;
      const int t = a.t0 + tile / tiles_per_t, b = (tile % tiles_per_t) * ROWS + lane * 4;
      if (b < a.B) {
        float* yp = view_ptr(a.y, g, t, a.B, b);
        const float* ap = a.epilogue == kEpiReluBwd ? view_ptr(a.aux, g, t, a.B, b) : nullptr;
#pragma unroll
        for (int j = 0; j < OPT; ++j) {
          const int o = og * OPT + j;
          if (o < out_dim) {
#pragma unroll
            for (int q = 0; q < RPT / 4; ++q) {
              if (b + 128 * q >= a.B) continue;   // RPT = 8: the second quad of rows may lie past the last env
              float4 v = make_float4(acc[4 * q][j], acc[4 * q + 1][j], acc[4 * q + 2][j], acc[4 * q + 3][j]);
              float4* dst = reinterpret_cast<float4*>(yp + (long long)o * a.B + 128 * q);
              if (a.epilogue == kEpiRelu) {
                v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
              } else if (a.epilogue == kEpiAccum) {
                const float4 o4 = *dst;
                v.x += o4.x, v.y += o4.y, v.z += o4.z, v.w += o4.w;
              } else if (a.epilogue == kEpiReluBwd) {
                const float4 m4 = *reinterpret_cast<const float4*>(ap + (long long)o * a.B + 128 * q);
                v.x = m4.x > 0.f ? v.x : 0.f, v.y = m4.y > 0.f ? v.y : 0.f;
                v.z = m4.z > 0.f ? v.z : 0.f, v.w = m4.w > 0.f ? v.w : 0.f;
              }
              *dst = v;
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done with buffer c & 1 before chunk c + 2 overwrites it
  }
}

// ------------------------------------------------------------------------------------------------
// Fused GRU step: gh = h_prev W_hh^T + b_hh (register-tiled as dense_tile_kernel<4, 24>) with the gate maths in
// the epilogue, so the [rows x 3H] hidden projection never goes to HBM.  Warp w owns hidden units
// [u0 + 8 w, u0 + 8 w + 8); its 24 accumulator columns are (r, z, n) x 8 units.  One launch covers 64 units; hidden
// sizes up to 128 (the reference's iRDQN networks: 100) take one launch per slice of 64 units (u0 = 0, 64).
// ------------------------------------------------------------------------------------------------
struct GruStepArgs {
  View h_prev, h_out, gi, acts;   // gi: input projections at time t - back (view.t_off carries the shift)
  int t_mod_gi;                   // > 0: gi is a ring of that many time blocks (rollout), index taken modulo
  const float* w;
  long long w_agent_stride;
  int whh_off[D2D_MAX_AGENTS];
  int bhh_off[D2D_MAX_AGENTS];
  int H, B, t0, t1;
  int back, padded, first, store, t_episode0;
  int u0;                         // first hidden unit of this launch (H > 64: one launch per slice of 64 units)
};

__device__ __forceinline__ float4 sigmoid4(float4 v) {
  return make_float4(1.0f / (1.0f + expf(-v.x)), 1.0f / (1.0f + expf(-v.y)), 1.0f / (1.0f + expf(-v.z)),
                     1.0f / (1.0f + expf(-v.w)));
}

__global__ void __launch_bounds__(256) gru_step_kernel(const GruStepArgs a) {
  constexpr int RPT = 4, OPT = 24, ROWS = 32 * RPT, KC = kDenseKC, OUT_PAD = 8 * OPT;
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y, H = a.H;
  const int in_dim = a.first ? 0 : H;
  float* Ms = sm;                                        // [in_dim][OUT_PAD], column = warp * 24 + gate * 8 + unit
  float* bs = Ms + (size_t)max(in_dim, 1) * OUT_PAD;
  float* Xs = bs + OUT_PAD;                              // [2][KC][ROWS]
  const float* W = a.w + g * a.w_agent_stride + a.whh_off[g];
  const float* bias = a.w + g * a.w_agent_stride + a.bhh_off[g];
  for (int i = threadIdx.x; i < in_dim * OUT_PAD; i += 256) {
    const int k = i / OUT_PAD, col = i % OUT_PAD;
    const int u = a.u0 + (col / OPT) * 8 + (col % 8), gate = (col % OPT) / 8;
    Ms[i] = u < H ? W[(long long)(gate * H + u) * H + k] : 0.f;
  }
  for (int col = threadIdx.x; col < OUT_PAD; col += 256) {
    const int u = a.u0 + (col / OPT) * 8 + (col % 8), gate = (col % OPT) / 8;
    bs[col] = u < H ? bias[gate * H + u] : 0.f;
  }
  const int lane = threadIdx.x & 31, og = threadIdx.x >> 5;
  const int tiles_per_t = (a.B + ROWS - 1) / ROWS;
  const int n_tiles = (a.t1 - a.t0) * tiles_per_t;
  const int nkc = max(1, (in_dim + KC - 1) / KC);
  const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int total = my_tiles * nkc;
  const long long HB = (long long)H * a.B;

  auto issue = [&](int c) {
    const int tile = blockIdx.x + (c / nkc) * gridDim.x;
    const int kc = c % nkc;
    const int t = a.t0 + tile / tiles_per_t, b0 = (tile % tiles_per_t) * ROWS;
    float* dst = Xs + (size_t)(c & 1) * KC * ROWS;
    const float* src = view_ptr(a.h_prev, g, t, a.B, b0);
    for (int i = threadIdx.x; i < KC * (ROWS / 4); i += 256) {
      const int kk = i / (ROWS / 4), r4 = (i % (ROWS / 4)) * 4;
      const int k = kc * KC + kk;
      float* d = dst + kk * ROWS + r4;
      if (k < in_dim && b0 + r4 < a.B) cp_async16(d, src + (long long)k * a.B + r4);
      else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  if (total > 0) issue(0);
  float acc[RPT][OPT];
  for (int c = 0; c < total; ++c) {
    if (c + 1 < total) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int kc = c % nkc;
    if (kc == 0) {
#pragma unroll
      for (int j = 0; j < OPT; ++j) {
        const float bv = bs[og * OPT + j];
#pragma unroll
        for (int r = 0; r < RPT; ++r) acc[r][j] = bv;
      }
    }
    const float* xs = Xs + (size_t)(c & 1) * KC * ROWS + lane * RPT;
    const int kmax = min(KC, in_dim - kc * KC);
    const float* mrow = Ms + (size_t)kc * KC * OUT_PAD + og * OPT;
#pragma unroll 2
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 v = *reinterpret_cast<const float4*>(xs + kk * ROWS);
      const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < OPT / 4; ++q) {
        const float4 m = *reinterpret_cast<const float4*>(mrow + kk * OUT_PAD + 4 * q);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          acc[r][4 * q + 0] = fmaf(xv[r], m.x, acc[r][4 * q + 0]);
          acc[r][4 * q + 1] = fmaf(xv[r], m.y, acc[r][4 * q + 1]);
          acc[r][4 * q + 2] = fmaf(xv[r], m.z, acc[r][4 * q + 2]);
          acc[r][4 * q + 3] = fmaf(xv[r], m.w, acc[r][4 * q + 3]);
        }
      }
    }
    if (kc == nkc - 1) {
      const int tile = blockIdx.x + (c / nkc) * gridDim.x;
      const int t = a.t0 + tile / tiles_per_t, b = (tile % tiles_per_t) * ROWS + lane * RPT;
      if (b < a.B) {
        const bool exists = a.padded || (t - a.back >= a.t_episode0);
        const float* hp = a.first ? nullptr : view_ptr(a.h_prev, g, t, a.B, b);
        float* ho = view_ptr(a.h_out, g, t, a.B, b);
        int tg = t + a.gi.t_off;
        if (a.t_mod_gi > 0) tg = ((tg % a.t_mod_gi) + a.t_mod_gi) % a.t_mod_gi;
        const float* gi = a.gi.p + (long long)tg * a.gi.t_stride + (long long)a.gi.f_off[g] * a.B + b;
        float* ac = a.store ? view_ptr(a.acts, g, t, a.B, b) : nullptr;
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          const int u = a.u0 + og * 8 + uu;
          if (u < H) {
            const long long uB = (long long)u * a.B;
            const float4 hpv = hp ? *reinterpret_cast<const float4*>(hp + uB) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 hn = hpv;
            if (exists) {
              const float4 gir = *reinterpret_cast<const float4*>(gi + uB);
              const float4 giz = *reinterpret_cast<const float4*>(gi + HB + uB);
              const float4 gin = *reinterpret_cast<const float4*>(gi + 2 * HB + uB);
              const float4 ghr = make_float4(acc[0][uu], acc[1][uu], acc[2][uu], acc[3][uu]);
              const float4 ghz = make_float4(acc[0][8 + uu], acc[1][8 + uu], acc[2][8 + uu], acc[3][8 + uu]);
              const float4 ghn = make_float4(acc[0][16 + uu], acc[1][16 + uu], acc[2][16 + uu], acc[3][16 + uu]);
              const float4 r = sigmoid4(make_float4(gir.x + ghr.x, gir.y + ghr.y, gir.z + ghr.z, gir.w + ghr.w));
              const float4 z = sigmoid4(make_float4(giz.x + ghz.x, giz.y + ghz.y, giz.z + ghz.z, giz.w + ghz.w));
              const float4 nn = make_float4(tanhf(gin.x + r.x * ghn.x), tanhf(gin.y + r.y * ghn.y),
                                            tanhf(gin.z + r.z * ghn.z), tanhf(gin.w + r.w * ghn.w));
              hn = make_float4((1.0f - z.x) * nn.x + z.x * hpv.x, (1.0f - z.y) * nn.y + z.y * hpv.y,
                               (1.0f - z.z) * nn.z + z.z * hpv.z, (1.0f - z.w) * nn.w + z.w * hpv.w);
              if (ac) {
                *reinterpret_cast<float4*>(ac + uB) = r;
                *reinterpret_cast<float4*>(ac + HB + uB) = z;
                *reinterpret_cast<float4*>(ac + 2 * HB + uB) = nn;
                *reinterpret_cast<float4*>(ac + 3 * HB + uB) = ghn;
              }
            }
            *reinterpret_cast<float4*>(ho + uB) = hn;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[o][k] (+ db[o]) = sum over rows of dy[o][row] * x[k][row]   (partial sums per row strip)
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  View dy, x;
  float* partial;              // [N][n_strips][part_stride]
  long long part_stride;       // >= out_dim * max_in + out_dim
  int in_dim[D2D_MAX_AGENTS];
  int out_dim;
  int B, t0, t1;
  int with_bias;
  int x_planes;                // tensor-core path: bf16 planes of x that can be non-zero (0 = all three; 1 when x is
                               // exactly representable in bf16 -- integer observations -- so only p0 carries data)
};

constexpr int kWgRows = 32;  // rows staged per iteration

template <int TO, int TK>    // each of the 16 x 16 threads owns outputs o = ty + 16 i (i < TO), k = tx + 16 j (j < TK)
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y, strip = blockIdx.x, n_strips = gridDim.x;
  const int K = a.in_dim[g], O = a.out_dim;
  constexpr int LD = kWgRows + 1;
  float* dys = sm;                 // [16 * TO][LD]
  float* xs = sm + 16 * TO * LD;   // [16 * TK][LD]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TO][TK];
  float accb[TO];
#pragma unroll
  for (int i = 0; i < TO; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TK; ++j) acc[i][j] = 0.f;
  }
  const int chunks_per_t = (a.B + kWgRows - 1) / kWgRows;
  const long long n_chunks = (long long)(a.t1 - a.t0) * chunks_per_t;
  for (long long c = strip; c < n_chunks; c += n_strips) {
    const int t = a.t0 + (int)(c / chunks_per_t);
    const int b0 = (int)(c % chunks_per_t) * kWgRows;
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * TO * kWgRows; i += 256) {
      const int o = i / kWgRows, r = i % kWgRows;
      dys[o * LD + r] = (o < O && b0 + r < a.B) ? view_ptr(a.dy, g, t, a.B, b0 + r)[(long long)o * a.B] : 0.f;
    }
    for (int i = threadIdx.x; i < 16 * TK * kWgRows; i += 256) {
      const int k = i / kWgRows, r = i % kWgRows;
      xs[k * LD + r] = (k < K && b0 + r < a.B) ? view_ptr(a.x, g, t, a.B, b0 + r)[(long long)k * a.B] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kWgRows; ++r) {
      float dv[TO], xv[TK];
#pragma unroll
      for (int i = 0; i < TO; ++i) dv[i] = dys[(ty + 16 * i) * LD + r];
#pragma unroll
      for (int j = 0; j < TK; ++j) xv[j] = xs[(tx + 16 * j) * LD + r];
#pragma unroll
      for (int i = 0; i < TO; ++i) {
        accb[i] += dv[i];
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
      }
    }
  }
  float* out = a.partial + ((long long)g * n_strips + strip) * a.part_stride;
#pragma unroll
  for (int i = 0; i < TO; ++i) {
    const int o = ty + 16 * i;
    if (o < O) {
#pragma unroll
      for (int j = 0; j < TK; ++j) {
        const int k = tx + 16 * j;
        if (k < K) out[(long long)o * K + k] = acc[i][j];
      }
      if (a.with_bias && tx == 0) out[(long long)O * K + o] = accb[i];
    }
  }
}

// grads[g][w_off + i] += sum_strips partial[g][s][i]  (fixed order: deterministic)
struct WreduceArgs {
  const float* partial;
  long long part_stride;
  int n_strips;
  float* grads;
  long long g_agent_stride;
  int w_off[D2D_MAX_AGENTS];
  int b_off[D2D_MAX_AGENTS];
  int in_dim[D2D_MAX_AGENTS];
  int out_dim;
  int with_bias;
};

__global__ void wreduce_kernel(const WreduceArgs a) {
  const int g = blockIdx.y;
  const int nw = a.out_dim * a.in_dim[g];
  const int n = nw + (a.with_bias ? a.out_dim : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    const float* p = a.partial + (long long)g * a.n_strips * a.part_stride + i;
    for (int k = 0; k < a.n_strips; ++k) s += p[(long long)k * a.part_stride];
    float* dst = a.grads + g * a.g_agent_stride + (i < nw ? a.w_off[g] + i : a.b_off[g] + (i - nw));
    *dst += s;
  }
}

}  // namespace d2d
