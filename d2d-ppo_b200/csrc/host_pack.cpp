// Host side of d2d_env_step_host: the reference's action array u8 [B][N][C] (0/1 per channel, what a caller of
// CombinatorialEnv.step(actions) holds: combinatorial_env.py:127) packed into the device layout -- one channel bitmask
// per (device, env), [N][B] -- BEFORE it crosses PCIe.  At C = 8 that is 8x fewer bytes on the link that bounds the
// host-buffer step (48 B of actions per env-step against 55 GB/s: DESIGN.md section 3).  Plain C++ (no CUDA): a small
// persistent thread pool and an AVX2 kernel (32 action bytes -> 4 mask bytes per compare + movemask), chosen at run time;
// the scalar loop covers every other CPU and channel count.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define D2D_X86 1
#endif

#include "../../include/d2d_b200.h"

namespace d2d {
void set_error(const char* fmt, ...);

namespace {

// Fork-join pool: run(f) calls f(part, parts) on `parts` threads (the caller is part 0) and returns when all are done.
// Created once and never destroyed (detached workers die with the process): no destructor order issues at unload.
class HostPool {
 public:
  explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
    for (int i = 1; i < n_; ++i) std::thread([this, i] { worker(i); }).detach();
  }
  int size() const { return n_; }
  void run(const std::function<void(int, int)>& f) {
    if (n_ == 1) {
      f(0, 1);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f, pending_ = n_ - 1, ++gen_;
    }
    cv_job_.notify_all();
    f(0, n_);
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  void worker(int id) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int, int)>* f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_job_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_, f = job_;
      }
      (*f)(id, n_);
      std::lock_guard<std::mutex> lk(m_);
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
  const int n_;
  std::mutex m_;
  std::condition_variable cv_job_, cv_done_;
  const std::function<void(int, int)>* job_ = nullptr;
  unsigned long long gen_ = 0;
  int pending_ = 0;
};

int g_threads = 0;            // 0: not chosen yet
HostPool* g_pool = nullptr;
std::mutex g_pool_mutex;

int default_threads() {
  unsigned hw = std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = (unsigned)CPU_COUNT(&set);   // respects a rank's NUMA binding
#endif
  if (hw == 0) hw = 1;
  // all CPUs, at most 16: packing alone is fastest with 12 of a 16-CPU box's CPUs (profiles/r02_u_time_host_pack.log),
  // but inside the host-buffer step 16 threads win (0.69 against 0.76 ms per step, profiles/r02_u_ab_host_pack.log)
  return (int)(hw > 16 ? 16 : hw);
}

HostPool* pool() {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  if (g_threads == 0) g_threads = default_threads();
  if (!g_pool || g_pool->size() != g_threads) g_pool = new HostPool(g_threads);   // a resized pool leaks the old one
  return g_pool;
}

template <typename MaskT>
void pack_scalar(const uint8_t* src, MaskT* dst, long long B, int N, int C, long long b0, long long b1) {
  for (long long b = b0; b < b1; ++b)
    for (int k = 0; k < N; ++k) {
      const uint8_t* a = src + ((size_t)b * N + k) * C;
      uint32_t m = 0;
      for (int c = 0; c < C; ++c) m |= (uint32_t)(a[c] != 0) << c;
      dst[(size_t)k * B + b] = (MaskT)m;
    }
}

#ifdef D2D_X86
// C == 8: one (env, device) pair is 8 bytes.  For a device k, the pairs of 8 consecutive envs (stride N x 8 bytes) are
// gathered into two 256-bit vectors; compare + movemask gives their 8 mask bytes, ONE 64-bit store into row k.  The
// N x 8 pairs of an env group are one contiguous 8 N x 8-byte window of the input, so every input line is read once.
__attribute__((target("avx2"))) void pack8_avx2(const uint8_t* src, uint8_t* dst, long long B, int N, long long b0,
                                                 long long b1) {
  const __m256i zero = _mm256_setzero_si256();
  const size_t stride = (size_t)N * 8;
  long long b = b0;
  for (; b + 8 <= b1; b += 8) {
    const uint8_t* base = src + (size_t)b * stride;
    for (int k = 0; k < N; ++k) {
      const uint8_t* q = base + (size_t)k * 8;
      long long w[8];
      for (int i = 0; i < 8; ++i) std::memcpy(&w[i], q + i * stride, 8);
      const __m256i v0 = _mm256_set_epi64x(w[3], w[2], w[1], w[0]);
      const __m256i v1 = _mm256_set_epi64x(w[7], w[6], w[5], w[4]);
      const uint64_t lo = (uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(v0, zero));
      const uint64_t hi = (uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(v1, zero));
      const uint64_t m = lo | (hi << 32);
      std::memcpy(dst + (size_t)k * B + b, &m, 8);
    }
  }
  for (; b < b1; ++b)
    for (int k = 0; k < N; ++k) {
      const uint8_t* a = src + ((size_t)b * N + k) * 8;
      uint32_t m = 0;
      for (int c = 0; c < 8; ++c) m |= (uint32_t)(a[c] != 0) << c;
      dst[(size_t)k * B + b] = (uint8_t)m;
    }
}
#endif

}  // namespace

// true where the vector kernel applies (8 channels, one mask byte, AVX2): only there does packing on the host outrun the
// PCIe copy of the unpacked bytes, so d2d_env_step_host takes the host path only then
bool host_pack_is_fast(int C, int mask_bytes) {
#ifdef D2D_X86
  return C == 8 && mask_bytes == 1 && __builtin_cpu_supports("avx2");
#else
  (void)C, (void)mask_bytes;
  return false;
#endif
}

// src u8 [B][N][C] -> dst masks [N][B] of mask_bytes (1 / 2 / 4) each; all threads of the pool, split by env ranges
void host_pack_actions(const uint8_t* src, void* dst, long long B, int N, int C, int mask_bytes) {
  HostPool* pl = pool();
  const bool avx2 = host_pack_is_fast(C, mask_bytes);
  const std::function<void(int, int)> job = [&](int part, int parts) {
    // whole cache lines of every output row per thread: env ranges in multiples of 64
    const long long per = ((B + parts - 1) / parts + 63) / 64 * 64;
    const long long b0 = (long long)part * per < B ? (long long)part * per : B;
    const long long b1 = b0 + per < B ? b0 + per : B;
    if (b0 >= b1) return;
#ifdef D2D_X86
    if (avx2) return pack8_avx2(src, reinterpret_cast<uint8_t*>(dst), B, N, b0, b1);
#endif
    if (mask_bytes == 1) pack_scalar(src, reinterpret_cast<uint8_t*>(dst), B, N, C, b0, b1);
    else if (mask_bytes == 2) pack_scalar(src, reinterpret_cast<uint16_t*>(dst), B, N, C, b0, b1);
    else pack_scalar(src, reinterpret_cast<uint32_t*>(dst), B, N, C, b0, b1);
  };
  pl->run(job);
}

int host_threads() {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  if (g_threads == 0) g_threads = default_threads();
  return g_threads;
}

}  // namespace d2d

extern "C" int d2d_set_host_threads(int n) {
  if (n < 0 || n > 256) {
    d2d::set_error("d2d_set_host_threads: %d not in 0..256", n);
    return D2D_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lk(d2d::g_pool_mutex);
  d2d::g_threads = n == 0 ? d2d::default_threads() : n;
  return D2D_OK;
}

extern "C" int d2d_get_host_threads(void) { return d2d::host_threads(); }

extern "C" int d2d_pack_actions_host(const uint8_t* actions_bnc, void* packed, int n_envs, int n_agents,
                                     int n_channels) {
  if (!actions_bnc || !packed || n_envs < 1 || n_agents < 1 || n_channels < 1 || n_channels > D2D_MAX_CHANNELS) {
    d2d::set_error("d2d_pack_actions_host: bad argument");
    return D2D_ERR_INVALID;
  }
  const int mask_bytes = n_channels <= 8 ? 1 : (n_channels <= 16 ? 2 : 4);
  d2d::host_pack_actions(actions_bnc, packed, n_envs, n_agents, n_channels, mask_bytes);
  return D2D_OK;
}
