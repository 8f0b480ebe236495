// Internal host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d3roma_b200.h"

namespace d3 {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define D3_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::d3::set_error(D3_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,        \
                             cudaGetErrorString(_e));                                         \
  } while (0)

#define D3_REQUIRE(cond, ...)                                                   \
  do {                                                                          \
    if (!(cond)) return ::d3::set_error(D3_ERR_INVALID, __VA_ARGS__);           \
  } while (0)

#define D3_LAUNCH_CHECK()                                                                     \
  do {                                                                                        \
    ::d3::count_launch();                                                                     \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return ::d3::set_error(D3_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,           \
                             cudaGetErrorString(_e));                                         \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace d3
