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

// num_chs == 1 (every pixel-space config): purely elementwise - four maps values per thread, 16-byte loads and stores,
// no index division; the arithmetic (and its rounding sequence) is that of denormalize_disp_kernel.
__global__ void __launch_bounds__(256) denormalize_disp_vec4_kernel(const float4* __restrict__ y, int64_t n4, float bound,
                                                                   float gamma, float t, float s,
                                                                   float4* __restrict__ out) {
  const float inv_gamma = 1.0f / gamma;
  auto f = [&](float x) {
    float v = __fadd_rn(__fdiv_rn(x, s), t);
    if (gamma != 1.0f) v = powf(v, inv_gamma);
    return __fadd_rn(0.f, __fmul_rn(v, bound));
  };
  D3_GRID_STRIDE(i, n4) {
    const float4 a = __ldcs(y + i);
    out[i] = make_float4(f(a.x), f(a.y), f(a.z), f(a.w));
  }
}

// Normalizer "piecewise" mode (utils_d3roma/utils.py:198-229): the disparity is split over up to 3 channels, channel c
// covering the range [sum_{j<c} bound_j, sum_{j<=c} bound_j):
//   ch0 = min(x, b0) / b0, res0 = (ch0 < 1) ? 0 : x - b0;  ch1 = min(res0, b1) / b1, res1 = (ch1 < 1) ? 0 : res0 - b1; ...
//   out_c = (ch_c ** gamma_c - t) * s        and back (utils.py:275-293):  sum_c (y_c / s + t) ** (1 / gamma_c) * b_c
struct PiecewiseArgs {
  float b[3], g[3];
  int n;
};

__global__ void normalize_piecewise_kernel(const float* __restrict__ disp, int64_t N, int64_t HW, const PiecewiseArgs a,
                                           float t, float s, float* __restrict__ out) {
  float total = 0.f;
  for (int c = 0; c < a.n; ++c) total = __fadd_rn(total, a.b[c]);
  D3_GRID_STRIDE(i, N * HW) {
    const int64_t n = i / HW, p = i - n * HW;
    float x = fminf(disp[i], total);
    for (int c = 0; c < a.n; ++c) {
      const float ch = __fdiv_rn(fminf(x, a.b[c]), a.b[c]);
      x = (ch < 1.0f) ? 0.f : __fsub_rn(x, a.b[c]);
      const float y = (a.g[c] != 1.0f) ? powf(ch, a.g[c]) : ch;
      out[(n * a.n + c) * HW + p] = __fmul_rn(__fsub_rn(y, t), s);
    }
  }
}

__global__ void denormalize_piecewise_kernel(const float* __restrict__ y, int64_t NR, int64_t HW, const PiecewiseArgs a,
                                             float t, float s, float* __restrict__ out) {
  D3_GRID_STRIDE(i, NR * HW) {
    const int64_t n = i / HW, p = i - n * HW;
    float z = 0.f;
    for (int c = 0; c < a.n; ++c) {
      float v = __fadd_rn(__fdiv_rn(y[(n * a.n + c) * HW + p], s), t);
      if (a.g[c] != 1.0f) v = powf(v, 1.0f / a.g[c]);
      const float term = __fmul_rn(v, a.b[c]);
      z = (c == 0) ? term : __fadd_rn(z, term);
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

// planes [N][H][W] -> [N][H+pt+pb][W+pl+pr], zero border (InputPadder.pad_zero, utils.py:317-319: the training mask)
__global__ void pad_zero_kernel(const float* __restrict__ in, int64_t N, int H, int W, int pl, int pr, int pt, int pb,
                                float* __restrict__ out) {
  const int Ho = H + pt + pb, Wo = W + pl + pr;
  D3_GRID_STRIDE(i, N * Ho * Wo) {
    const int x = (int)(i % Wo) - pl;
    const int64_t r = i / Wo;
    const int y = (int)(r % Ho) - pt;
    const int64_t n = r / Ho;
    out[i] = (x >= 0 && x < W && y >= 0 && y < H) ? in[(n * H + y) * W + x] : 0.f;
  }
}

// train.py:175-178: x <- clamp(x, -r, r); mask <- mask * (|x| < r), mask [B][1][HW] broadcast over the C channels of x
__global__ void clamp_target_kernel(float* __restrict__ x, float* __restrict__ mask, int64_t B, int C, int64_t HW, float r) {
  D3_GRID_STRIDE(i, B * HW) {
    const int64_t b = i / HW, p = i - b * HW;
    float m = mask[i];
    for (int c = 0; c < C; ++c) {
      float* px = x + (b * C + c) * HW + p;
      const float v = fminf(fmaxf(*px, -r), r);
      *px = v;
      m *= (fabsf(v) < r) ? 1.f : 0.f;
    }
    mask[i] = m;
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
  pdl_prologue();
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
  if (num_chs == 1 && (NR * HW) % 4 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    const int64_t n4 = NR * HW / 4;
    denormalize_disp_vec4_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(y), n4, bound, gamma, t, s, reinterpret_cast<float4*>(out));
    D3_LAUNCH_CHECK();
    return D3_OK;
  }
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

extern "C" int d3_pad_zero(const float* in, int64_t planes, int H, int W, int left, int right, int top, int bottom,
                           float* out, void* stream) {
  D3_REQUIRE(in && out && H > 0 && W > 0 && left >= 0 && right >= 0 && top >= 0 && bottom >= 0, "pad_zero: bad arguments");
  const int64_t n = planes * (H + top + bottom) * (W + left + right);
  if (n == 0) return D3_OK;
  pad_zero_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, planes, H, W, left, right, top, bottom, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_clamp_target(float* x, float* mask, int64_t B, int C, int64_t HW, float range, void* stream) {
  D3_REQUIRE(x && mask && B > 0 && C > 0 && HW > 0 && range > 0.f, "clamp_target: bad arguments");
  clamp_target_kernel<<<ew_blocks(B * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mask, B, C, HW, range);
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
  D3_CHECK_CUDA(launch_pdl(clamp_into_kernel, dim3(ew_blocks(B * chw)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                           in, B, chw, out_bstride, out_off, lo, hi, out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_normalize_piecewise(const float* disp, int64_t N, int64_t HW, int num_chs, const float* bounds,
                                      const float* gammas, float t, float s, float* out, void* stream) {
  using namespace d3;
  D3_REQUIRE(disp && out && bounds && gammas, "normalize_piecewise: null pointer");
  D3_REQUIRE(num_chs >= 1 && num_chs <= 3, "normalize_piecewise: num_chs must be 1..3 (got %d)", num_chs);
  PiecewiseArgs a{};
  a.n = num_chs;
  for (int c = 0; c < num_chs; ++c) { a.b[c] = bounds[c]; a.g[c] = gammas[c]; }
  if (N * HW == 0) return D3_OK;
  normalize_piecewise_kernel<<<ew_blocks(N * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, N, HW, a, t, s, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_denormalize_piecewise(const float* y, int64_t NR, int64_t HW, int num_chs, const float* bounds,
                                        const float* gammas, float t, float s, float* out, void* stream) {
  using namespace d3;
  D3_REQUIRE(y && out && bounds && gammas, "denormalize_piecewise: null pointer");
  D3_REQUIRE(num_chs >= 1 && num_chs <= 3, "denormalize_piecewise: num_chs must be 1..3 (got %d)", num_chs);
  PiecewiseArgs a{};
  a.n = num_chs;
  for (int c = 0; c < num_chs; ++c) { a.b[c] = bounds[c]; a.g[c] = gammas[c]; }
  if (NR * HW == 0) return D3_OK;
  denormalize_piecewise_kernel<<<ew_blocks(NR * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, NR, HW, a, t, s, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

namespace d3 {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// Depth map -> colour (RGB) camera frame: back-project every valid depth pixel with the depth intrinsics, move it into
// the colour frame, project with the colour intrinsics and splat the ORIGINAL depth value at the rounded pixel
// (DepthCamera.transform_depth_to_rgb_frame / transform_cropped_depth_to_rgb_frame,
// /root/reference/utils_d3roma/camera.py:610-701).  The reference is numpy: float64 matrix products, then
// `depth_map[v, u] = values` - a fancy-index assignment in which, for colliding targets, the LAST source pixel in
// row-major order wins; with "aggressive_fill" four such assignments run one after the other on the
// (floor, floor), (floor, ceil), (ceil, floor), (ceil, ceil) corners, later ones overwriting earlier ones.
// Both orders are reproduced exactly with one 32-bit atomicMax per splat on the key (fill index, source index), followed
// by a gather of the winning source's depth (fused with the output crop).  Arithmetic is fp64 like numpy's.
// ------------------------------------------------------------------------------------------------------------------
struct ReprojArgs {
  double Kinv[9];   // inverse depth intrinsics, row-major
  double T[12];     // colour <- depth rigid transform, row-major 3x4
  double Kc[9];     // colour intrinsics, row-major
  int H, W;         // source size
  int CH, CW;       // canvas (colour image) size
  int aggressive;
};

__global__ void reproject_scatter_kernel(const float* __restrict__ depth, const ReprojArgs a,
                                         unsigned int* __restrict__ keys) {
  const int64_t n = (int64_t)a.H * a.W;
  D3_GRID_STRIDE(i, n) {
    const float dz = depth[i];
    const int y = (int)(i / a.W), x = (int)(i - (int64_t)y * a.W);
    const double fx = (double)(float)x, fy = (double)(float)y;   // numpy builds the pixel grid in float32
    // (Kinv @ [x, y, 1]) * depth
    const double px = (a.Kinv[0] * fx + a.Kinv[1] * fy + a.Kinv[2]) * (double)dz;
    const double py = (a.Kinv[3] * fx + a.Kinv[4] * fy + a.Kinv[5]) * (double)dz;
    const double pz = (a.Kinv[6] * fx + a.Kinv[7] * fy + a.Kinv[8]) * (double)dz;
    if (!(pz > 0.0)) continue;
    const double cx = a.T[0] * px + a.T[1] * py + a.T[2] * pz + a.T[3];
    const double cy = a.T[4] * px + a.T[5] * py + a.T[6] * pz + a.T[7];
    const double cz = a.T[8] * px + a.T[9] * py + a.T[10] * pz + a.T[11];
    const double qx = a.Kc[0] * cx + a.Kc[1] * cy + a.Kc[2] * cz;
    const double qy = a.Kc[3] * cx + a.Kc[4] * cy + a.Kc[5] * cz;
    const double qz = a.Kc[6] * cx + a.Kc[7] * cy + a.Kc[8] * cz;
    const double u = qx / qz, v = qy / qz;
    const int nfill = a.aggressive ? 4 : 1;
    for (int f = 0; f < nfill; ++f) {
      double uu, vv;
      if (a.aggressive) {
        uu = (f & 2) ? ceil(u) : floor(u);     // fills: (left, up), (left, bottom), (right, up), (right, bottom)
        vv = (f & 1) ? ceil(v) : floor(v);
      } else {
        uu = rint(u);                           // np.round: half to even
        vv = rint(v);
      }
      // astype(np.uint32) wraps negative values to huge ones (rejected by the `< W` test); -0.0 becomes 0
      if (!(uu >= 0.0 && uu < (double)a.CW && vv >= 0.0 && vv < (double)a.CH)) continue;
      const unsigned int key = ((unsigned int)(f + 1) << 28) | (unsigned int)i;
      atomicMax(keys + (size_t)(int)vv * a.CW + (int)uu, key);
    }
  }
}

__global__ void reproject_gather_kernel(const float* __restrict__ depth, const unsigned int* __restrict__ keys, int CW,
                                        int oy, int ox, int OH, int OW, float* __restrict__ out) {
  D3_GRID_STRIDE(i, (int64_t)OH * OW) {
    const int y = (int)(i / OW), x = (int)(i - (int64_t)y * OW);
    const unsigned int k = keys[(size_t)(y + oy) * CW + (x + ox)];
    out[i] = k ? depth[k & 0x0FFFFFFFu] : 0.f;
  }
}

}  // namespace
}  // namespace d3

extern "C" size_t d3_depth_to_rgb_frame_workspace_bytes(int canvas_h, int canvas_w) {
  return (size_t)canvas_h * canvas_w * sizeof(unsigned int);
}

extern "C" int d3_depth_to_rgb_frame(const float* depth, int H, int W, const double* Kinv_depth, const double* T_cl,
                                     const double* K_color, int canvas_h, int canvas_w, int aggressive_fill, int crop_y,
                                     int crop_x, int out_h, int out_w, void* workspace, float* out, void* stream) {
  using namespace d3;
  D3_REQUIRE(depth && Kinv_depth && T_cl && K_color && workspace && out, "depth_to_rgb_frame: null pointer");
  D3_REQUIRE(H > 0 && W > 0 && (int64_t)H * W < (1 << 28), "depth_to_rgb_frame: bad source size %dx%d", H, W);
  D3_REQUIRE(canvas_h > 0 && canvas_w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + out_h <= canvas_h &&
                 crop_x + out_w <= canvas_w && out_h > 0 && out_w > 0,
             "depth_to_rgb_frame: crop window outside the %dx%d canvas", canvas_h, canvas_w);
  ReprojArgs a;
  for (int i = 0; i < 9; ++i) { a.Kinv[i] = Kinv_depth[i]; a.Kc[i] = K_color[i]; }
  for (int i = 0; i < 12; ++i) a.T[i] = T_cl[i];
  a.H = H; a.W = W; a.CH = canvas_h; a.CW = canvas_w; a.aggressive = aggressive_fill ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  D3_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)canvas_h * canvas_w * sizeof(unsigned int), st));
  reproject_scatter_kernel<<<ew_blocks((int64_t)H * W), 256, 0, st>>>(depth, a, static_cast<unsigned int*>(workspace));
  D3_LAUNCH_CHECK();
  reproject_gather_kernel<<<ew_blocks((int64_t)out_h * out_w), 256, 0, st>>>(
      depth, static_cast<const unsigned int*>(workspace), canvas_w, crop_y, crop_x, out_h, out_w, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
