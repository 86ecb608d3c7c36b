// Shared host/device helpers for libd2d_b200.so (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d2d_b200.h"

namespace d2d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a grid-stride kernel over `n` items: whole waves of the 148-SM part
inline int grid_for(long long n, int block, int max_blocks_per_sm = 8) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  long long need = (n + block - 1) / block;
  long long cap = (long long)sms * max_blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace d2d

#define D2D_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      d2d::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return D2D_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define D2D_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      d2d::set_error(__VA_ARGS__);      \
      return D2D_ERR_INVALID;           \
    }                                   \
  } while (0)

#define D2D_LAUNCHED()                  \
  do {                                  \
    d2d::count_launch();                \
    D2D_CUDA(cudaGetLastError());       \
  } while (0)
