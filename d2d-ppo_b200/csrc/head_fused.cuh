// Network head in one pass: out = W2 relu(W1 h + b1) + b2 per row (reference: the `layers` Sequential of RNN,
// algorithms/d2d_ppo.py:36-41,54: Linear(H, H) -> ReLU -> Linear(H, out)), FP32 on the CUDA cores.
//
// Replaces two grouped-GEMM launches and the y1 round trip through HBM on the rollout path (y1 is still written when
// the caller needs it for the backward pass).  A thread owns two rows (envs b and b + 128 of a 256-row tile) with
// both hidden vectors in registers; W1 is staged transposed in shared memory so that eight consecutive outputs of one
// input are two broadcast LDS.128 feeding 16 FMAs; the second layer is accumulated on the fly from each group of
// eight first-layer outputs, so y1 never leaves registers.  4 H (H + O) flops per thread against H^2 / 4 LDS.128:
// bound by the FMA pipe, not by shared memory.
#pragma once
#include "learner_kernels.cuh"

namespace d2d {

struct HeadFusedArgs {
  View h;      // [.. H ..] in
  View y1;     // [.. H ..] out (p == nullptr: not stored)
  View out;    // [.. O ..] out (pre-activation)
  const float* w;
  long long w_agent_stride;
  int w1_off[D2D_MAX_AGENTS], b1_off[D2D_MAX_AGENTS], w2_off[D2D_MAX_AGENTS], b2_off[D2D_MAX_AGENTS];
  int O, B, t0, t1;
};

constexpr int kHeadThreads = 128;

template <int H, int OMAX, bool KEEP_Y1>   // KEEP_Y1: the backward pass needs relu(W1 h + b1) (training direction)
__global__ void __launch_bounds__(kHeadThreads) head_fused_kernel(const HeadFusedArgs a) {
  static_assert(H % 8 == 0 && H <= 64, "hidden size");
  __shared__ __align__(16) float w1s[H * H];        // [k][j] = W1[j][k]
  __shared__ __align__(16) float w2s[H * OMAX];     // [j][o] = W2[o][j]
  __shared__ __align__(16) float b1s[H];
  __shared__ float b2s[OMAX];
  const int g = blockIdx.y, tid = threadIdx.x, O = a.O;
  const float* base = a.w + g * a.w_agent_stride;
  for (int i = tid; i < H * H; i += kHeadThreads) {
    const int j = i / H, k = i % H;                 // coalesced read of W1[j][k]
    w1s[k * H + j] = base[a.w1_off[g] + i];
  }
  for (int i = tid; i < H * OMAX; i += kHeadThreads) {
    const int j = i / OMAX, o = i % OMAX;
    w2s[i] = o < O ? base[a.w2_off[g] + o * H + j] : 0.f;
  }
  for (int i = tid; i < H; i += kHeadThreads) b1s[i] = base[a.b1_off[g] + i];
  for (int i = tid; i < OMAX; i += kHeadThreads) b2s[i] = i < O ? base[a.b2_off[g] + i] : 0.f;
  __syncthreads();

  const int tiles_per_t = (a.B + 2 * kHeadThreads - 1) / (2 * kHeadThreads);
  const int n_tiles = (a.t1 - a.t0) * tiles_per_t;
  const long long Bl = a.B;
  for (int p = blockIdx.x; p < n_tiles; p += gridDim.x) {
    const int t = a.t0 + p / tiles_per_t;
    const int b0 = (p % tiles_per_t) * (2 * kHeadThreads) + tid, b1 = b0 + kHeadThreads;
    const bool ok0 = b0 < a.B, ok1 = b1 < a.B;
    const float* hp0 = view_ptr(a.h, g, t, a.B, ok0 ? b0 : 0);
    const float* hp1 = view_ptr(a.h, g, t, a.B, ok1 ? b1 : 0);
    float h0[H], h1[H];
#pragma unroll
    for (int k = 0; k < H; ++k) h0[k] = ok0 ? hp0[k * Bl] : 0.f, h1[k] = ok1 ? hp1[k * Bl] : 0.f;
    float o0[OMAX], o1[OMAX];
#pragma unroll
    for (int o = 0; o < OMAX; ++o) o0[o] = o1[o] = b2s[o];
    float* yp0 = KEEP_Y1 ? view_ptr(a.y1, g, t, a.B, ok0 ? b0 : 0) : nullptr;
    float* yp1 = KEEP_Y1 ? view_ptr(a.y1, g, t, a.B, ok1 ? b1 : 0) : nullptr;
#pragma unroll 1
    for (int jc = 0; jc < H / 8; ++jc) {
      float a0[8], a1[8];
      {
        const float4 ba = *reinterpret_cast<const float4*>(b1s + jc * 8);
        const float4 bb = *reinterpret_cast<const float4*>(b1s + jc * 8 + 4);
        a0[0] = a1[0] = ba.x, a0[1] = a1[1] = ba.y, a0[2] = a1[2] = ba.z, a0[3] = a1[3] = ba.w;
        a0[4] = a1[4] = bb.x, a0[5] = a1[5] = bb.y, a0[6] = a1[6] = bb.z, a0[7] = a1[7] = bb.w;
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float4 wa = *reinterpret_cast<const float4*>(w1s + k * H + jc * 8);
        const float4 wb = *reinterpret_cast<const float4*>(w1s + k * H + jc * 8 + 4);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) a0[i] = fmaf(w[i], h0[k], a0[i]), a1[i] = fmaf(w[i], h1[k], a1[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = jc * 8 + i;
        a0[i] = fmaxf(a0[i], 0.f), a1[i] = fmaxf(a1[i], 0.f);
        if (KEEP_Y1 && ok0) yp0[j * Bl] = a0[i];
        if (KEEP_Y1 && ok1) yp1[j * Bl] = a1[i];
#pragma unroll
        for (int o = 0; o < OMAX; ++o) {
          const float w2 = w2s[j * OMAX + o];
          o0[o] = fmaf(w2, a0[i], o0[o]), o1[o] = fmaf(w2, a1[i], o1[o]);
        }
      }
    }
    float* op0 = view_ptr(a.out, g, t, a.B, ok0 ? b0 : 0);
    float* op1 = view_ptr(a.out, g, t, a.B, ok1 ? b1 : 0);
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
      if (o < O) {
        if (ok0) op0[o * Bl] = o0[o];
        if (ok1) op1[o * Bl] = o1[o];
      }
    }
  }
}

}  // namespace d2d
