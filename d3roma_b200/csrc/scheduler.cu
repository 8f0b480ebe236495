// Fused DDIM / DDPM scheduler updates (one elementwise pass instead of ~12-15 ATen launches).
//
// Replaces the tensor arithmetic of MyDDIMScheduler.step (/root/reference/core/scheduler_ddim.py:423-489) and
// MyDDPMScheduler.step (/root/reference/core/scheduler_ddpm.py:460-518).  The scalar coefficients are computed by
// the host-side scheduler classes with the same fp32 tensor expressions as the reference; the kernels perform
// exactly the reference's sequence of fp32 roundings (explicit __f*_rn intrinsics, no FMA contraction), so the
// result is bit-identical to the reference's fp32 CPU path.  "Dynamic thresholding" with sample_max_value = 1.0
// is clamp(-1, 1) / 1 - the per-image quantile sort is not computed.
//
// Algorithmic bytes: 4 * n * 4 B (read model_output, sample; write prev_sample, pred_original_sample)
// (+ n * 4 B for the variance noise of DDPM / eta > 0).
#include "common.cuh"
#include "internal.h"
#include "sched_math.cuh"

namespace d3 {
namespace {

__global__ void __launch_bounds__(256) ddim_step_kernel(const float* __restrict__ mo, const float* __restrict__ xs,
                                                        const float* __restrict__ noise, d3_ddim_coeffs c, int64_t n,
                                                        float* __restrict__ prev, float* __restrict__ orig,
                                                        float* __restrict__ prev_mean, float* __restrict__ dir_out) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mo[i], s = xs[i];
    float x0, eps;
    split_x0_eps(c.prediction_type, m, s, c.sqrt_alpha_t, c.sqrt_beta_t, x0, eps);
    x0 = clip_x0(x0, c.clip_mode, c.clip_range);
    if (c.use_clipped_model_output) eps = __fdiv_rn(__fsub_rn(s, __fmul_rn(c.sqrt_alpha_t, x0)), c.sqrt_beta_t);
    const float dir = __fmul_rn(c.dir_coeff, eps);
    const float p = __fadd_rn(__fmul_rn(c.sqrt_alpha_prev, x0), dir);
    orig[i] = x0;
    if (prev_mean) prev_mean[i] = p;
    if (dir_out) dir_out[i] = dir;
    prev[i] = noise ? __fadd_rn(p, __fmul_rn(c.std_dev, noise[i])) : p;
  }
}

__global__ void __launch_bounds__(256) ddpm_step_kernel(const float* __restrict__ mo, const float* __restrict__ xs,
                                                        const float* __restrict__ noise, d3_ddpm_coeffs c, int64_t n,
                                                        float* __restrict__ prev, float* __restrict__ orig,
                                                        float* __restrict__ prev_mean) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mo[i], s = xs[i];
    float x0, eps;
    split_x0_eps(c.prediction_type, m, s, c.sqrt_alpha_t, c.sqrt_beta_t, x0, eps);
    x0 = clip_x0(x0, c.clip_mode, c.clip_range);
    const float mean = __fadd_rn(__fmul_rn(c.x0_coeff, x0), __fmul_rn(c.xt_coeff, s));
    orig[i] = x0;
    if (prev_mean) prev_mean[i] = mean;
    prev[i] = noise ? __fadd_rn(mean, __fmul_rn(c.sigma, noise[i])) : mean;
  }
}

// out = a*x + y  or  out = a*x + b*y (+ c*z): the recombination after a guidance perturbation of x0
__global__ void __launch_bounds__(256) lincomb_kernel(float a, const float* __restrict__ x, float b,
                                                      const float* __restrict__ y, float c,
                                                      const float* __restrict__ z, int64_t n,
                                                      float* __restrict__ out) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float r = __fmul_rn(a, x[i]);
    if (y) r = __fadd_rn(r, (b == 1.0f) ? y[i] : __fmul_rn(b, y[i]));
    if (z) r = __fadd_rn(r, __fmul_rn(c, z[i]));
    out[i] = r;
  }
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace d3

extern "C" int d3_sched_ddim_step(const float* model_output, const float* sample, const float* variance_noise,
                                  const d3_ddim_coeffs* c, int64_t n, float* prev_sample, float* pred_original,
                                  float* prev_mean, float* dir_out, void* stream) {
  using namespace d3;
  D3_REQUIRE(model_output && sample && c && prev_sample && pred_original && n >= 0, "ddim_step: bad arguments");
  D3_REQUIRE(c->prediction_type >= 0 && c->prediction_type <= 2, "ddim_step: bad prediction_type %d",
             c->prediction_type);
  if (n == 0) return D3_OK;
  D3_CHECK_CUDA(launch_pdl(ddim_step_kernel, dim3(ew_blocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), model_output, sample, variance_noise, *c, n, prev_sample, pred_original, prev_mean, dir_out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_sched_ddpm_step(const float* model_output, const float* sample, const float* variance_noise,
                                  const d3_ddpm_coeffs* c, int64_t n, float* prev_sample, float* pred_original,
                                  float* prev_mean, void* stream) {
  using namespace d3;
  D3_REQUIRE(model_output && sample && c && prev_sample && pred_original && n >= 0, "ddpm_step: bad arguments");
  D3_REQUIRE(c->prediction_type >= 0 && c->prediction_type <= 2, "ddpm_step: bad prediction_type %d",
             c->prediction_type);
  if (n == 0) return D3_OK;
  D3_CHECK_CUDA(launch_pdl(ddpm_step_kernel, dim3(ew_blocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), model_output, sample, variance_noise, *c, n, prev_sample, pred_original, prev_mean));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_lincomb(float a, const float* x, float b, const float* y, float c, const float* z, int64_t n,
                          float* out, void* stream) {
  using namespace d3;
  D3_REQUIRE(x && out && n >= 0, "lincomb: bad arguments");
  if (n == 0) return D3_OK;
  D3_CHECK_CUDA(launch_pdl(lincomb_kernel, dim3(ew_blocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), a, x, b, y, c, z, n, out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}
