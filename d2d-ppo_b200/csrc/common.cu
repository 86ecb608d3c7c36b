// Error string, ABI version and launch counter of libd2d_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace d2d {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace d2d

extern "C" {
const char* d2d_last_error(void) { return d2d::g_err; }
int d2d_abi_version(void) { return D2D_ABI_VERSION; }
uint64_t d2d_launch_count(void) { return d2d::g_launches.load(std::memory_order_relaxed); }
}
