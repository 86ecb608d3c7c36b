// Philox4x32-10 (Salmon et al., SC'11) and the 16-bit-lane draw helpers of the "philox" RNG mode.
// Counter layout: (env_global_index, timestep, device | purpose << 16, block); key = 64-bit seed.
// Arrival uniforms share a call between four devices, the single-channel env's 1-bit draws between eight
// (env_common.cuh: ArrivalWords, LaneWords).
// The numpy restatement used by the parity tests is oracle/philox_np.py.
#pragma once
#include <stdint.h>

namespace d2d {

constexpr uint32_t kPurposeSwitch = 0u, kPurposeArrival = 1u, kPurposePolicy = 2u;
constexpr uint32_t kEnvLevelDevice = 0xFFFFu;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// bit c of the result = (16-bit lane c < thr(c)), c in [0, n_lanes)
template <typename ThrFn>
__device__ __forceinline__ uint32_t philox_lane_mask(uint32_t env, uint32_t t, uint32_t dev_purpose, int n_lanes,
                                                     uint32_t k0, uint32_t k1, ThrFn thr) {
  uint32_t m = 0;
  for (int blk = 0; blk * 8 < n_lanes; ++blk) {
    const uint4 r = philox4x32_10(env, t, dev_purpose, (uint32_t)blk, k0, k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const int c = blk * 8 + l;
      if (c < n_lanes) {
        const uint32_t u = (w[l >> 1] >> (16 * (l & 1))) & 0xFFFFu;
        m |= (uint32_t)(u < thr(c)) << c;
      }
    }
  }
  return m;
}

}  // namespace d2d
