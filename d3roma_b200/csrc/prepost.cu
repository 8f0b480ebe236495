// Pre/post-processing either side of the denoising loop, fused into single elementwise passes:
//   depth -> disparity -> normalised disparity   (inference.py:128-131 + Normalizer.normalize, utils.py:171-196)
//   normalised disparity -> disparity             (Normalizer.denormalize "average" mode, utils.py:252-273)
//   disparity -> depth                            (inference.py:246-249)
//   replicate pad / crop                          (InputPadder.pad / unpad, utils.py:313-325)
// (reference root: /root/reference/utils_d3roma/utils.py, /root/reference/inference.py)
// The reference does these with boolean-mask indexing (sync + gather/scatter) and F.pad copies; here each is one
// coalesced pass with the same fp32 operation order (bit-identical for gamma == 1).
// Algorithmic bytes: 2 * H * W * 4 per map (one read, one write).
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

#define D3_GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void depth_to_disp_kernel(const float* __restrict__ depth, float fxb, int64_t n, float* __restrict__ disp) {
  D3_GRID_STRIDE(i, n) {
    // the reference evaluates `python_float / tensor`, which torch computes as tensor.reciprocal() * scalar
    // (Tensor.__rtruediv__): two roundings, reproduced here so the result is bit-identical.
    const float d = depth[i];
    disp[i] = d > 0.f ? __fmul_rn(__frcp_rn(d), fxb) : 0.f;
  }
}

// in [N][HW] raw disparity -> out [N][num_chs][HW] normalised: ((min(x, bound) / bound) ** gamma - t) * s
__global__ void normalize_disp_kernel(const float* __restrict__ disp, int64_t N, int64_t HW, int num_chs, float bound,
                                      float gamma, float t, float s, float* __restrict__ out) {
  D3_GRID_STRIDE(i, N * HW) {
    const int64_t n = i / HW, p = i - n * HW;
    float y = __fdiv_rn(fminf(disp[i], bound), bound);
    if (gamma != 1.0f) y = powf(y, gamma);
    const float v = __fmul_rn(__fsub_rn(y, t), s);
    for (int c = 0; c < num_chs; ++c) out[(n * num_chs + c) * HW + p] = v;
  }
}

// in [N][R*num_chs][HW] -> out [N][R][HW]: sum_c (y_c / s + t) ** (1/gamma) * (bound / num_chs)
__global__ void denormalize_disp_kernel(const float* __restrict__ y, int64_t NR, int64_t HW, int num_chs, float bound,
                                        float gamma, float t, float s, float* __restrict__ out) {
  const float per = bound / (float)num_chs;
  const float inv_gamma = 1.0f / gamma;
  D3_GRID_STRIDE(i, NR * HW) {
    const int64_t n = i / HW, p = i - n * HW;
    float z = 0.f;
    for (int c = 0; c < num_chs; ++c) {
      float v = __fadd_rn(__fdiv_rn(y[(n * num_chs + c) * HW + p], s), t);
      if (gamma != 1.0f) v = powf(v, inv_gamma);
      z = __fadd_rn(z, __fmul_rn(v, per));
    }
    out[i] = z;
  }
}

__global__ void disp_to_depth_kernel(const float* __restrict__ disp, float fxb, int64_t n, float* __restrict__ depth) {
  D3_GRID_STRIDE(i, n) {
    const float d = disp[i];
    depth[i] = d > 0.f ? __fdiv_rn(fxb, d) : 0.f;
  }
}

// planes [N][H][W] -> [N][H+pt+pb][W+pl+pr], replicate border
__global__ void pad_replicate_kernel(const float* __restrict__ in, int64_t N, int H, int W, int pl, int pr, int pt,
                                     int pb, float* __restrict__ out) {
  const int Ho = H + pt + pb, Wo = W + pl + pr;
  D3_GRID_STRIDE(i, N * Ho * Wo) {
    const int x = (int)(i % Wo);
    const int64_t r = i / Wo;
    const int y = (int)(r % Ho);
    const int64_t n = r / Ho;
    const int sy = min(max(y - pt, 0), H - 1), sx = min(max(x - pl, 0), W - 1);
    out[i] = in[(n * H + sy) * W + sx];
  }
}

__global__ void crop_kernel(const float* __restrict__ in, int64_t N, int Hi, int Wi, int top, int left, int H, int W,
                            float* __restrict__ out) {
  D3_GRID_STRIDE(i, N * H * W) {
    const int x = (int)(i % W);
    const int64_t r = i / W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    out[i] = in[(n * Hi + y + top) * Wi + x + left];
  }
}

// clamp to [-1, 1] (custom_pipelines.py:196) into a slot of a wider channel stack: out[b][c_off + c]
__global__ void clamp_into_kernel(const float* __restrict__ in, int64_t B, int64_t CHW, int64_t out_bstride,
                                  int64_t out_off, float lo, float hi, float* __restrict__ out) {
  D3_GRID_STRIDE(i, B * CHW) {
    const int64_t b = i / CHW, r = i - b * CHW;
    out[b * out_bstride + out_off + r] = fminf(fmaxf(in[i], lo), hi);
  }
}

}  // namespace
}  // namespace d3

using namespace d3;

extern "C" int d3_depth_to_disp(const float* depth, float fxb, int64_t n, float* disp, void* stream) {
  D3_REQUIRE(depth && disp && n >= 0, "depth_to_disp: bad arguments");
  if (n == 0) return D3_OK;
  depth_to_disp_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(depth, fxb, n, disp);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_normalize_disp(const float* disp, int64_t N, int64_t HW, int num_chs, float bound, float gamma,
                                 float t, float s, float* out, void* stream) {
  D3_REQUIRE(disp && out && num_chs >= 1 && bound > 0.f && gamma > 0.f, "normalize_disp: bad arguments");
  if (N * HW == 0) return D3_OK;
  normalize_disp_kernel<<<ew_blocks(N * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, N, HW, num_chs, bound,
                                                                                         gamma, t, s, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_denormalize_disp(const float* y, int64_t NR, int64_t HW, int num_chs, float bound, float gamma,
                                   float t, float s, float* out, void* stream) {
  D3_REQUIRE(y && out && num_chs >= 1 && bound > 0.f && gamma > 0.f && s != 0.f, "denormalize_disp: bad arguments");
  if (NR * HW == 0) return D3_OK;
  denormalize_disp_kernel<<<ew_blocks(NR * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, NR, HW, num_chs,
                                                                                           bound, gamma, t, s, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_disp_to_depth(const float* disp, float fxb, int64_t n, float* depth, void* stream) {
  D3_REQUIRE(disp && depth && n >= 0, "disp_to_depth: bad arguments");
  if (n == 0) return D3_OK;
  disp_to_depth_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, fxb, n, depth);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_pad_replicate(const float* in, int64_t planes, int H, int W, int left, int right, int top,
                                int bottom, float* out, void* stream) {
  D3_REQUIRE(in && out && H > 0 && W > 0 && left >= 0 && right >= 0 && top >= 0 && bottom >= 0,
             "pad_replicate: bad arguments");
  const int64_t n = planes * (H + top + bottom) * (W + left + right);
  if (n == 0) return D3_OK;
  pad_replicate_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, planes, H, W, left, right, top,
                                                                                   bottom, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_crop(const float* in, int64_t planes, int Hin, int Win, int top, int left, int H, int W, float* out,
                       void* stream) {
  D3_REQUIRE(in && out && top >= 0 && left >= 0 && top + H <= Hin && left + W <= Win, "crop: bad arguments");
  const int64_t n = planes * H * W;
  if (n == 0) return D3_OK;
  crop_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, planes, Hin, Win, top, left, H, W, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_clamp_into(const float* in, int64_t B, int64_t chw, int64_t out_bstride, int64_t out_off, float lo,
                             float hi, float* out, void* stream) {
  D3_REQUIRE(in && out && B >= 0 && chw >= 0, "clamp_into: bad arguments");
  if (B * chw == 0) return D3_OK;
  clamp_into_kernel<<<ew_blocks(B * chw), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, B, chw, out_bstride,
                                                                                      out_off, lo, hi, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
