// Evaluation-side post-processing on the device: the reference moves predictions to the host
// (.cpu().numpy(), evaluate.py:122-124) and loops over samples in numpy; here each is a per-sample reduction kernel.
//   d3_disp_metrics        compute_errors(space="disp")            /root/reference/utils_d3roma/utils.py:554-653
//   d3_scale_shift_lsq     compute_scale_and_shift (SSI alignment)  utils_d3roma/utils.py:488-528
//   d3_masked_minmax, d3_normalize_ssi, d3_affine_rows              Normalizer ssi=True branch, utils.py:171-182, 231-251
// One 1024-thread block per sample (B <= a few hundred, 230 K pixels each): these are microsecond-scale kernels; the
// point is removing the D2H sync + numpy loop from the evaluation sweep, not bandwidth records.
#include <cfloat>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kRedThreads = 1024;

template <typename T, typename Op>
D3_DEVINL T block_reduce(T v, Op op, T* scratch) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  v = scratch[lane % (kRedThreads / 32)];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;  // every thread holds the block-wide result
}

struct SumD { D3_DEVINL double operator()(double a, double b) const { return a + b; } };
struct MinF { D3_DEVINL float operator()(float a, float b) const { return fminf(a, b); } };
struct MaxF { D3_DEVINL float operator()(float a, float b) const { return fmaxf(a, b); } };

// out[b] = {epe, d1, d2, d3, d5, a1, a2, a3, rmse, rel, mae}
__global__ void __launch_bounds__(kRedThreads) disp_metrics_kernel(const float* __restrict__ gt,
                                                                   const float* __restrict__ pred,
                                                                   const float* __restrict__ mask,
                                                                   const float* __restrict__ fxb, int fxb_n, int64_t HW,
                                                                   double* __restrict__ out) {
  __shared__ double sd[32];
  __shared__ float sf[32];
  const int b = blockIdx.x;
  const float* g = gt + (size_t)b * HW;
  const float* p = pred + (size_t)b * HW;
  const float* m = mask + (size_t)b * HW;
  // pass 1: number of valid pixels, min of the (masked) prediction, min of the ground truth
  double cnt = 0.0;
  float pmin = FLT_MAX, gmin = FLT_MAX;
  for (int64_t i = threadIdx.x; i < HW; i += kRedThreads) {
    if (m[i] != 0.f) {
      cnt += 1.0;
      pmin = fminf(pmin, p[i]);
      gmin = fminf(gmin, g[i]);
    }
  }
  cnt = block_reduce(cnt, SumD(), sd);
  pmin = block_reduce(pmin, MinF(), sf);
  gmin = block_reduce(gmin, MinF(), sf);
  const bool clip = pmin <= 0.f;                 // "you predicted zero disp values": clip to the smallest gt disparity
  const float scale = (float)(cnt / 480.0);      // sic: the reference scales the D-k thresholds by n_valid / 480 (utils.py:629)
  const float f = fxb_n > 0 ? fxb[fxb_n > 1 ? b : 0] : 1.0f;
  double s_epe = 0, s_d1 = 0, s_d2 = 0, s_d3 = 0, s_d5 = 0, s_a1 = 0, s_a2 = 0, s_a3 = 0, s_sq = 0, s_rel = 0, s_abs = 0;
  for (int64_t i = threadIdx.x; i < HW; i += kRedThreads) {
    if (m[i] != 0.f) {
      const float gd = g[i];
      float pd = p[i];
      if (clip) pd = fmaxf(pd, gmin);
      const float e = fabsf(gd - pd);
      s_epe += e;
      s_d1 += e > 1.0f * scale; s_d2 += e > 2.0f * scale; s_d3 += e > 3.0f * scale; s_d5 += e > 5.0f * scale;
      const float gz = __fdiv_rn(f, gd), pz = __fdiv_rn(f, pd);
      const float th = fmaxf(__fdiv_rn(gz, pz), __fdiv_rn(pz, gz));
      s_a1 += th < 1.05f; s_a2 += th < 1.10f; s_a3 += th < 1.25f;
      const float dz = gz - pz;
      s_sq += (double)(dz * dz);
      s_rel += (double)__fdiv_rn(fabsf(dz), gz);
      s_abs += fabsf(dz);
    }
  }
  double r[11] = {s_epe, s_d1, s_d2, s_d3, s_d5, s_a1, s_a2, s_a3, s_sq, s_rel, s_abs};
#pragma unroll
  for (int k = 0; k < 11; ++k) r[k] = block_reduce(r[k], SumD(), sd);
  if (threadIdx.x == 0) {
    double* o = out + (size_t)b * 11;
    for (int k = 0; k < 11; ++k) o[k] = r[k] / cnt;
    o[8] = sqrt(o[8]);
  }
}

// per row: least-squares (scale, shift) with prediction * s + t ~ target over masked pixels
__global__ void __launch_bounds__(kRedThreads) scale_shift_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ target,
                                                                  const float* __restrict__ mask, int64_t HW,
                                                                  float* __restrict__ out) {
  __shared__ double sd[32];
  const int b = blockIdx.x;
  const float* p = pred + (size_t)b * HW;
  const float* t = target + (size_t)b * HW;
  const float* m = mask + (size_t)b * HW;
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (int64_t i = threadIdx.x; i < HW; i += kRedThreads) {
    if (m[i] != 0.f) {
      const double pv = p[i], tv = t[i];
      a00 += pv * pv; a01 += pv; a11 += 1.0; b0 += pv * tv; b1 += tv;
    }
  }
  a00 = block_reduce(a00, SumD(), sd); a01 = block_reduce(a01, SumD(), sd); a11 = block_reduce(a11, SumD(), sd);
  b0 = block_reduce(b0, SumD(), sd); b1 = block_reduce(b1, SumD(), sd);
  if (threadIdx.x == 0) {
    double det = a00 * a11 - a01 * a01;
    if (!(det > 0.0)) det += 1e-4;  // utils.py:518-520
    out[2 * b] = (float)((a11 * b0 - a01 * b1) / det);
    out[2 * b + 1] = (float)((-a01 * b0 + a00 * b1) / det);
  }
}

// ---- RANSAC scale/shift (utils_d3roma/ransac.py:51-96; Normalizer.denormalize with safe_ssi, utils.py:240-243) ----
// one iteration = {fit on the N sampled pixels, inlier set of that model, refit on the inliers, global accuracy score,
// per-row keep-if-better}; the random permutation comes from the caller (torch.randperm, as the reference draws it).
__global__ void __launch_bounds__(kRedThreads) scale_shift_gather_kernel(const float* __restrict__ pred,
                                                                         const float* __restrict__ target,
                                                                         const float* __restrict__ mask,
                                                                         const int* __restrict__ idx, int N, int64_t HW,
                                                                         float* __restrict__ out) {
  __shared__ double sd[32];
  const int b = blockIdx.x;
  const float* p = pred + (size_t)b * HW;
  const float* t = target + (size_t)b * HW;
  const float* m = mask + (size_t)b * HW;
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (int i = threadIdx.x; i < N; i += kRedThreads) {
    const int j = idx[i];
    if (m[j] != 0.f) {
      const double pv = p[j], tv = t[j];
      a00 += pv * pv; a01 += pv; a11 += 1.0; b0 += pv * tv; b1 += tv;
    }
  }
  a00 = block_reduce(a00, SumD(), sd); a01 = block_reduce(a01, SumD(), sd); a11 = block_reduce(a11, SumD(), sd);
  b0 = block_reduce(b0, SumD(), sd); b1 = block_reduce(b1, SumD(), sd);
  if (threadIdx.x == 0) {
    double det = a00 * a11 - a01 * a01;
    if (!(det > 0.0)) det += 1e-4;
    out[2 * b] = (float)((a11 * b0 - a01 * b1) / det);
    out[2 * b + 1] = (float)((-a01 * b0 + a00 * b1) / det);
  }
}

// thr[b][i] = ((target - (pred * s + t))^2 < T) & mask ; n_inlier[b] = sum(thr)      (ransac.py:80-81, 86)
__global__ void __launch_bounds__(kRedThreads) ransac_inliers_kernel(const float* __restrict__ pred,
                                                                     const float* __restrict__ target,
                                                                     const float* __restrict__ mask,
                                                                     const float* __restrict__ model, float T, int64_t HW,
                                                                     float* __restrict__ thr, float* __restrict__ n_inlier) {
  __shared__ double sd[32];
  const int b = blockIdx.x;
  const float s = model[2 * b], t = model[2 * b + 1];
  double cnt = 0.0;
  for (int64_t i = threadIdx.x; i < HW; i += kRedThreads) {
    const size_t o = (size_t)b * HW + i;
    const float x_ = __fadd_rn(__fmul_rn(pred[o], s), t);
    const float e = __fsub_rn(target[o], x_);
    const bool in = (__fmul_rn(e, e) < T) && (mask[o] != 0.f);
    thr[o] = in ? 1.0f : 0.f;
    cnt += in ? 1.0 : 0.0;
  }
  cnt = block_reduce(cnt, SumD(), sd);
  if (threadIdx.x == 0) n_inlier[b] = (float)cnt;
}

// count of elements (ALL rows, all pixels - the reference's mean has no dim) with max(Y/X'', X''/Y) < 1.25  (ransac.py:19-21)
__global__ void __launch_bounds__(256) ransac_score_kernel(const float* __restrict__ pred,
                                                           const float* __restrict__ target,
                                                           const float* __restrict__ model, int rows, int64_t HW,
                                                           unsigned long long* __restrict__ count) {
  unsigned int local = 0;
  const int64_t total = (int64_t)rows * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const float x = __fadd_rn(__fmul_rn(pred[i], model[2 * b]), model[2 * b + 1]);
    const float y = target[i];
    const float r0 = __fdiv_rn(y, x), r1 = __fdiv_rn(x, y);
    // torch.maximum propagates NaN, and NaN < 1.25 is false
    if (!(r0 != r0) && !(r1 != r1) && fmaxf(r0, r1) < 1.25f) ++local;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
}

// keep the refit where it has more than D inliers and a better (global) score   (ransac.py:87-92)
__global__ void ransac_select_kernel(int rows, const float* __restrict__ n_inlier, float D,
                                     const unsigned long long* __restrict__ count, double total,
                                     const float* __restrict__ better, float* __restrict__ best_num,
                                     float* __restrict__ best_err, float* __restrict__ best_fit) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= rows) return;
  const float mean_acc = (float)((double)*count / total);
  const float this_err = 1.0f / mean_acc;   // inf when nothing is within 1.25x
  if (n_inlier[b] > D && this_err < best_err[b]) {
    best_num[b] = n_inlier[b];
    best_err[b] = this_err;
    best_fit[2 * b] = better[2 * b];
    best_fit[2 * b + 1] = better[2 * b + 1];
  }
}

__global__ void __launch_bounds__(kRedThreads) masked_minmax_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ mask, int64_t HW,
                                                                    float* __restrict__ out) {
  __shared__ float sf[32];
  const int b = blockIdx.x;
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int64_t i = threadIdx.x; i < HW; i += kRedThreads) {
    if (mask == nullptr || mask[(size_t)b * HW + i] != 0.f) {
      const float v = x[(size_t)b * HW + i];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
  }
  lo = block_reduce(lo, MinF(), sf);
  hi = block_reduce(hi, MaxF(), sf);
  if (threadIdx.x == 0) {
    out[2 * b] = lo;
    out[2 * b + 1] = hi;
  }
}

// y = mask ? (clamp((x - low) / (up - low), 0, 1) - t) * s : 0      (utils.py:171-180)
__global__ void normalize_ssi_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                     const float* __restrict__ lowup, int64_t N, int64_t HW, float t, float s,
                                     float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * HW; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW;
    const float low = lowup[2 * n], up = lowup[2 * n + 1];
    float y = 0.f;
    if (mask == nullptr || mask[i] != 0.f) {
      const float v = __fdiv_rn(__fsub_rn(x[i], low), __fsub_rn(up, low));
      y = __fmul_rn(__fsub_rn(fminf(fmaxf(v, 0.f), 1.f), t), s);
    }
    out[i] = y;
  }
}

// out[r][i] = y[r][i] * st[r][0] + st[r][1]     (utils.py:250-251 `y * s + t`)
__global__ void affine_rows_kernel(const float* __restrict__ y, const float* __restrict__ st, int64_t R, int64_t HW,
                                   float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R * HW; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / HW;
    out[i] = __fadd_rn(__fmul_rn(y[i], st[2 * r]), st[2 * r + 1]);
  }
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace d3

using namespace d3;

extern "C" int d3_disp_metrics(const float* gt_disp, const float* pred_disp, const float* mask, const float* fxb,
                               int fxb_n, int B, int64_t HW, double* out, void* stream) {
  D3_REQUIRE(gt_disp && pred_disp && mask && out && B >= 0 && HW > 0 && (fxb_n == 0 || fxb), "disp_metrics: bad arguments");
  D3_REQUIRE(fxb_n == 0 || fxb_n == 1 || fxb_n == B, "disp_metrics: fxb must have 0, 1 or B entries");
  if (B == 0) return D3_OK;
  disp_metrics_kernel<<<B, kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(gt_disp, pred_disp, mask, fxb, fxb_n, HW,
                                                                              out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_scale_shift_lsq(const float* pred, const float* target, const float* mask, int rows, int64_t HW,
                                  float* scale_shift, void* stream) {
  D3_REQUIRE(pred && target && mask && scale_shift && rows >= 0 && HW > 0, "scale_shift: bad arguments");
  if (rows == 0) return D3_OK;
  scale_shift_kernel<<<rows, kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, mask, HW, scale_shift);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_masked_minmax(const float* x, const float* mask, int rows, int64_t HW, float* low_up, void* stream) {
  D3_REQUIRE(x && low_up && rows >= 0 && HW > 0, "masked_minmax: bad arguments");
  if (rows == 0) return D3_OK;
  masked_minmax_kernel<<<rows, kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, mask, HW, low_up);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_normalize_ssi(const float* x, const float* mask, const float* low_up, int rows, int64_t HW, float t,
                                float s, float* out, void* stream) {
  D3_REQUIRE(x && low_up && out && rows >= 0 && HW > 0, "normalize_ssi: bad arguments");
  if (rows == 0) return D3_OK;
  normalize_ssi_kernel<<<ew_blocks(rows * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mask, low_up, rows, HW, t,
                                                                                           s, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_affine_rows(const float* y, const float* scale_shift, int rows, int64_t HW, float* out, void* stream) {
  D3_REQUIRE(y && scale_shift && out && rows >= 0 && HW > 0, "affine_rows: bad arguments");
  if (rows == 0) return D3_OK;
  affine_rows_kernel<<<ew_blocks(rows * HW), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, scale_shift, rows, HW, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" size_t d3_ransac_workspace_bytes(int rows, int64_t HW) {
  // inlier mask [rows][HW] + maybe-model, refit, inlier counts [rows] x 5 floats + one 64-bit counter
  return ((size_t)rows * HW + (size_t)rows * 5) * sizeof(float) + 16;
}

extern "C" int d3_ransac_iteration(const float* pred, const float* target, const float* mask, const int* perm, int n_sample,
                                   int rows, int64_t HW, float T, float min_inliers, void* workspace, float* best_fit,
                                   float* best_err, float* best_num, void* stream) {
  D3_REQUIRE(pred && target && mask && perm && workspace && best_fit && best_err && best_num, "ransac: null pointer");
  D3_REQUIRE(rows > 0 && HW > 0 && n_sample > 0 && n_sample <= HW, "ransac: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* count = static_cast<unsigned long long*>(workspace);
  float* thr = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16);
  float* maybe = thr + (size_t)rows * HW;
  float* better = maybe + 2 * rows;
  float* n_in = better + 2 * rows;
  D3_CHECK_CUDA(cudaMemsetAsync(count, 0, 16, st));
  scale_shift_gather_kernel<<<rows, kRedThreads, 0, st>>>(pred, target, mask, perm, n_sample, HW, maybe);
  D3_LAUNCH_CHECK();
  ransac_inliers_kernel<<<rows, kRedThreads, 0, st>>>(pred, target, mask, maybe, T, HW, thr, n_in);
  D3_LAUNCH_CHECK();
  scale_shift_kernel<<<rows, kRedThreads, 0, st>>>(pred, target, thr, HW, better);
  D3_LAUNCH_CHECK();
  ransac_score_kernel<<<ew_blocks((int64_t)rows * HW), 256, 0, st>>>(pred, target, better, rows, HW, count);
  D3_LAUNCH_CHECK();
  ransac_select_kernel<<<(rows + 127) / 128, 128, 0, st>>>(rows, n_in, min_inliers, count, (double)rows * (double)HW,
                                                          better, best_num, best_err, best_fit);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
