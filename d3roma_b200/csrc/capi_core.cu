// Library-level C-ABI entry points: error text, version, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "internal.h"

namespace d3 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static std::atomic<int> g_pdl{-1};   // -1: not decided yet (D3_PDL, default off); 0 / 1: set by d3_set_pdl or the env

bool pdl_enabled() {
  int on = g_pdl.load(std::memory_order_relaxed);
  if (on < 0) {
    const char* e = getenv("D3_PDL");
    on = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    g_pdl.store(on, std::memory_order_relaxed);
  }
  return on != 0;
}

cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[std::make_pair(dev, func)];
  if (bytes > have) {
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    have = bytes;
  }
  return cudaSuccess;
}

}  // namespace d3

extern "C" {

const char* d3_last_error(void) { return d3::g_err; }
int d3_version(void) { return 100; }
int64_t d3_launch_count(void) { return d3::g_launches.load(std::memory_order_relaxed); }
void d3_reset_launch_count(void) { d3::g_launches.store(0, std::memory_order_relaxed); }
int d3_set_pdl(int on) {
  const int prev = d3::pdl_enabled() ? 1 : 0;
  d3::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

}  // extern "C"
