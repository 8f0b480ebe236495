// Internal host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d3roma_b200.h"

namespace d3 {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();  // SM count of the CURRENT device
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size): the attribute is per device, so
// the cache is too (one process may drive several GPUs).
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);

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

// fp32 reference-precision variants (fp32_path.cu), reached through the same C-ABI entry points with D3_DTYPE_F32
int conv2d_f32(const d3_conv_params* p, cudaStream_t stream);
int groupnorm_f32(const void* x1, int C1, int p1, const double* stats1, const void* x2, int C2, int p2,
                  const double* stats2, const float* gamma, const float* beta, int B, int HW, int groups, float eps,
                  int apply_silu, void* out, int pout, cudaStream_t stream);
int attention_f32(const void* qkv, int pitch, int B, int N, int C, void* out, int opitch, cudaStream_t stream);
int conv_out_f32(const void* x, int pitch, const double* stats, const float* gamma, const float* beta, const float* w,
                 const float* bias, int B, int H, int W, int dc, float eps, float* out, cudaStream_t stream);
int im2col_f32(const float* sample, int Cs, const float* cond, int Cc, int B, int H, int W, void* out,
               cudaStream_t stream);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): the kernels of the denoising loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that the next kernel's blocks become resident and run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch, index arithmetic) while the previous kernel drains;
// every such kernel calls pdl_prologue() before its first global-memory access (griddepcontrol.wait returns only when
// the prerequisite grid has completed and its writes are visible).  D3_PDL=0 launches them the plain way.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const bool on = pdl_enabled();
  cfg.attrs = on ? attr : nullptr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace d3
