// ensemble_depths (utils_d3roma/utils.py:28-120; evaluate.py:117-120 `config.ensemble`): n <= 32 affine-invariant maps
// of one sample are aligned by per-map scale / shift (found by scipy's BFGS on the host, which calls the objective
// below once per evaluation), reduced per pixel (median or mean), and rescaled to [0, 1].
//
//   d3_ensemble_objective   one evaluation of the closure of utils.py:64-87 on the device:
//                             v_i = x_i * s_i + t_i            (a rounded product, then a rounded sum - as torch does)
//                             out[0] = sum over pairs i < j and pixels of (v_i - v_j)^2   (torch.combinations order)
//                             out[1], out[2] = min / max over pixels of pred = median_i v_i (lower median, as
//                                              torch.median) or mean_i v_i
//                           partial sums per block in fp64, combined in a fixed order: the value is deterministic, which
//                           the finite differences of BFGS need.
//   d3_ensemble_reduce      the prediction of utils.py:99-118 with the final s, t: aligned = median / mean, uncertainty =
//                           median absolute deviation / unbiased std, both rescaled by the aligned map's min / max.
// HBM-bound streaming kernels: every map is read exactly once per call, coalesced (consecutive threads = consecutive
// pixels of each map), 4 * n bytes per pixel algorithmic.
#include "common.cuh"
#include "internal.h"

namespace d3 {

constexpr int kEnsMaxN = 32;
constexpr int kEnsThreads = 256;
constexpr int kEnsMaxBlocks = 1184;   // 8 per SM

struct EnsArgs {
  const float* images;  // [n][P]
  int n;
  int64_t P;
  int median;           // 1: median / MAD, 0: mean / std
  float s[kEnsMaxN], t[kEnsMaxN];
};

// The kernels are instantiated for every map count up to 8 (NT = n: loops unroll, the per-pixel values live in
// registers) and once generically (NT = 0: runtime n <= 32, values in local memory).
template <int NT>
D3_DEVINL int count_of(int n) { return NT ? NT : n; }

// lower median of v[0..n) (torch.median: the element of rank (n - 1) / 2), by rank counting - no data movement
template <int NT>
D3_DEVINL float lower_median(const float* v, int n_rt) {
  const int n = count_of<NT>(n_rt);
  const int k = (n - 1) >> 1;
  float med = v[0];
#pragma unroll
  for (int i = 0; i < n; ++i) {
    int rank = 0;
#pragma unroll
    for (int j = 0; j < n; ++j) rank += (v[j] < v[i]) || (v[j] == v[i] && j < i);
    if (rank == k) med = v[i];
  }
  return med;
}

template <int NT>
D3_DEVINL float mean_of(const float* v, int n_rt) {
  const int n = count_of<NT>(n_rt);
  float a = v[0];
#pragma unroll
  for (int i = 1; i < n; ++i) a = __fadd_rn(a, v[i]);   // torch's outer-dimension reduction adds the rows in order
  return __fdiv_rn(a, (float)n);
}

// block reduction of (sum, min, max) -> partial[blockIdx.x * 3 ..]
D3_DEVINL void block_partials(double sum, float mn, float mx, double* __restrict__ partial) {
  __shared__ double s_sum[kEnsThreads / 32];
  __shared__ float s_mn[kEnsThreads / 32], s_mx[kEnsThreads / 32];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_sum[wid] = sum; s_mn[wid] = mn; s_mx[wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    float lo = INFINITY, hi = -INFINITY;
    for (int w = 0; w < kEnsThreads / 32; ++w) { a += s_sum[w]; lo = fminf(lo, s_mn[w]); hi = fmaxf(hi, s_mx[w]); }
    partial[blockIdx.x * 3 + 0] = a;
    partial[blockIdx.x * 3 + 1] = (double)lo;
    partial[blockIdx.x * 3 + 2] = (double)hi;
  }
}

template <int NT>
__global__ void __launch_bounds__(kEnsThreads) ensemble_objective_kernel(const EnsArgs a, double* __restrict__ partial) {
  float v[NT ? NT : kEnsMaxN];
  const int n = count_of<NT>(a.n);
  double sum = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t p = (int64_t)blockIdx.x * kEnsThreads + threadIdx.x; p < a.P; p += (int64_t)gridDim.x * kEnsThreads) {
#pragma unroll
    for (int i = 0; i < n; ++i) v[i] = __fadd_rn(__fmul_rn(__ldg(a.images + (size_t)i * a.P + p), a.s[i]), a.t[i]);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = i + 1; j < n; ++j) {
        const double d = (double)__fsub_rn(v[i], v[j]);    // the fp32 difference of the reference, squared and summed in fp64
        sum = fma(d, d, sum);
      }
    const float pred = a.median ? lower_median<NT>(v, n) : mean_of<NT>(v, n);
    mn = fminf(mn, pred);
    mx = fmaxf(mx, pred);
  }
  block_partials(sum, mn, mx, partial);
}

template <int NT>
__global__ void __launch_bounds__(kEnsThreads) ensemble_reduce_kernel(const EnsArgs a, float* __restrict__ aligned,
                                                                     float* __restrict__ uncertainty,
                                                                     double* __restrict__ partial) {
  float v[NT ? NT : kEnsMaxN], dev[NT ? NT : kEnsMaxN];
  const int n = count_of<NT>(a.n);
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t p = (int64_t)blockIdx.x * kEnsThreads + threadIdx.x; p < a.P; p += (int64_t)gridDim.x * kEnsThreads) {
#pragma unroll
    for (int i = 0; i < n; ++i) v[i] = __fadd_rn(__fmul_rn(__ldg(a.images + (size_t)i * a.P + p), a.s[i]), a.t[i]);
    float al, un;
    if (a.median) {
      al = lower_median<NT>(v, n);
#pragma unroll
      for (int i = 0; i < n; ++i) dev[i] = fabsf(__fsub_rn(v[i], al));
      un = lower_median<NT>(dev, n);
    } else {
      al = mean_of<NT>(v, n);
      // torch.std(dim=0): unbiased; two-pass in fp64 here (torch runs Welford's update: equal to within an ulp or two)
      double m = 0.0, q = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) m += (double)v[i];
      m /= (double)n;
#pragma unroll
      for (int i = 0; i < n; ++i) { const double d = (double)v[i] - m; q += d * d; }
      un = n > 1 ? (float)sqrt(q / (double)(n - 1)) : __int_as_float(0x7fc00000);
    }
    aligned[p] = al;
    uncertainty[p] = un;
    mn = fminf(mn, al);
    mx = fmaxf(mx, al);
  }
  block_partials(0.0, mn, mx, partial);
}

// one block: combine the per-block partials in index order -> out3 = {sum, min, max}
__global__ void __launch_bounds__(kEnsThreads) ensemble_finalize_kernel(const double* __restrict__ partial, int blocks,
                                                                       double* __restrict__ out3) {
  __shared__ double s_sum[kEnsThreads], s_mn[kEnsThreads], s_mx[kEnsThreads];
  double a = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int b = threadIdx.x; b < blocks; b += kEnsThreads) {
    a += partial[b * 3];
    lo = fmin(lo, partial[b * 3 + 1]);
    hi = fmax(hi, partial[b * 3 + 2]);
  }
  s_sum[threadIdx.x] = a; s_mn[threadIdx.x] = lo; s_mx[threadIdx.x] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kEnsThreads; ++i) { a += s_sum[i]; lo = fmin(lo, s_mn[i]); hi = fmax(hi, s_mx[i]); }
    out3[0] = a; out3[1] = lo; out3[2] = hi;
  }
}

// aligned = (aligned - min) / (max - min); uncertainty /= (max - min)   (utils.py:113-117, fp32 like the reference)
__global__ void __launch_bounds__(kEnsThreads) ensemble_rescale_kernel(float* __restrict__ aligned,
                                                                      float* __restrict__ uncertainty, int64_t P,
                                                                      const double* __restrict__ out3) {
  const float lo = (float)out3[1], range = __fsub_rn((float)out3[2], (float)out3[1]);
  for (int64_t p = (int64_t)blockIdx.x * kEnsThreads + threadIdx.x; p < P; p += (int64_t)gridDim.x * kEnsThreads) {
    aligned[p] = __fdiv_rn(__fsub_rn(aligned[p], lo), range);
    uncertainty[p] = __fdiv_rn(uncertainty[p], range);
  }
}

static int ens_blocks(int64_t P) {
  const int64_t want = (P + kEnsThreads - 1) / kEnsThreads;
  return (int)(want < kEnsMaxBlocks ? want : kEnsMaxBlocks);
}

static int fill_args(EnsArgs& a, const float* images, int n, int64_t P, const float* s, const float* t, int reduction) {
  if (!(images && s && t)) return set_error(D3_ERR_INVALID, "ensemble: null pointer");
  if (!(n >= 1 && n <= kEnsMaxN && P > 0)) return set_error(D3_ERR_INVALID, "ensemble: 1 <= n <= %d maps expected (got %d)", kEnsMaxN, n);
  if (!(reduction == 0 || reduction == 1)) return set_error(D3_ERR_INVALID, "ensemble: reduction must be 0 (mean) or 1 (median)");
  a.images = images; a.n = n; a.P = P; a.median = reduction;
  for (int i = 0; i < n; ++i) { a.s[i] = s[i]; a.t[i] = t[i]; }
  return D3_OK;
}

}  // namespace d3

using namespace d3;

extern "C" size_t d3_ensemble_workspace_bytes(void) { return (size_t)(kEnsMaxBlocks * 3 + 3) * sizeof(double); }

extern "C" int d3_ensemble_objective(const float* images, int n, int64_t P, const float* s_host, const float* t_host,
                                     int reduction, void* workspace, double* out3, void* stream) {
  EnsArgs a;
  if (int rc = fill_args(a, images, n, P, s_host, t_host, reduction)) return rc;
  D3_REQUIRE(workspace && out3, "ensemble_objective: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ens_blocks(P);
  double* partial = static_cast<double*>(workspace);
  switch (n) {
#define D3_ENS_CASE(N) case N: ensemble_objective_kernel<N><<<blocks, kEnsThreads, 0, st>>>(a, partial); break;
    D3_ENS_CASE(1) D3_ENS_CASE(2) D3_ENS_CASE(3) D3_ENS_CASE(4) D3_ENS_CASE(5) D3_ENS_CASE(6) D3_ENS_CASE(7) D3_ENS_CASE(8)
#undef D3_ENS_CASE
    default: ensemble_objective_kernel<0><<<blocks, kEnsThreads, 0, st>>>(a, partial);
  }
  D3_LAUNCH_CHECK();
  ensemble_finalize_kernel<<<1, kEnsThreads, 0, st>>>(partial, blocks, out3);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_ensemble_reduce(const float* images, int n, int64_t P, const float* s_host, const float* t_host,
                                  int reduction, void* workspace, float* aligned, float* uncertainty, void* stream) {
  EnsArgs a;
  if (int rc = fill_args(a, images, n, P, s_host, t_host, reduction)) return rc;
  D3_REQUIRE(workspace && aligned && uncertainty, "ensemble_reduce: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ens_blocks(P);
  double* partial = static_cast<double*>(workspace);
  double* out3 = partial + (size_t)kEnsMaxBlocks * 3;
  switch (n) {
#define D3_ENS_CASE(N) case N: ensemble_reduce_kernel<N><<<blocks, kEnsThreads, 0, st>>>(a, aligned, uncertainty, partial); break;
    D3_ENS_CASE(1) D3_ENS_CASE(2) D3_ENS_CASE(3) D3_ENS_CASE(4) D3_ENS_CASE(5) D3_ENS_CASE(6) D3_ENS_CASE(7) D3_ENS_CASE(8)
#undef D3_ENS_CASE
    default: ensemble_reduce_kernel<0><<<blocks, kEnsThreads, 0, st>>>(a, aligned, uncertainty, partial);
  }
  D3_LAUNCH_CHECK();
  ensemble_finalize_kernel<<<1, kEnsThreads, 0, st>>>(partial, blocks, out3);
  D3_LAUNCH_CHECK();
  ensemble_rescale_kernel<<<blocks, kEnsThreads, 0, st>>>(aligned, uncertainty, P, out3);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
