// GroupNorm (32 groups) + optional SiLU, applied as one bandwidth-bound pass over bf16 NHWC activations.
//
// Replaces aten::native_group_norm (+ aten::silu) of every ResnetBlock2D.norm1/norm2, Attention.group_norm and
// conv_norm_out (diffusers ResnetBlock2D built at /root/reference/train.py:788-810).
//
// Statistics are NOT computed here: every producer (the conv epilogue) accumulates, per sample, the sum and
// sum-of-squares of each 4-channel "quad" of its bf16-rounded output in fp64 ([B][C/4][2]).  All group widths in
// this UNet (4, 8, 12, 16, 24, 32 channels) are multiples of 4, and so are the split points of every skip
// concatenation, so a group's moments are the sum of its quads' moments — including groups that straddle the two
// source tensors of a virtual concat (C = 384 = 256 + 128, C = 768 = 512 + 256).  The kernel reads up
// to two sources and writes ONE dense normalised tensor, i.e. the torch.cat of the up-blocks costs nothing extra.
//
// Algorithmic bytes: 2 * N * sizeof(bf16) per layer (one read, one write).
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kGnThreads = 256;

constexpr int kMaxC = 1024;  // nvec = C/8 <= 128 threads per pixel

struct GnArgs {
  const uint16_t* x1;
  const uint16_t* x2;
  const double* st1;
  const double* st2;
  const float* gamma;
  const float* beta;
  uint16_t* out;
  int C1, C2, p1, p2, pout;
  int HW;        // pixels per sample
  int groups;
  float eps;
  int silu;
  int pix_per_block;
  // dual mode (single source only): a second normalisation of the SAME input with its own affine parameters and
  // grouping, written to a second tensor - the up-path copy of a skip tensor (see d3_groupnorm_silu_dual)
  const float* gamma2;
  const float* beta2;
  uint16_t* out2;
  int pout2, groups2, silu2;
};

template <int DT, bool DUAL = false, int U = 8>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const GnArgs a) {
  pdl_prologue();
  // blockDim.x = nvec * k (nvec = C/8 16-byte vectors per pixel): every thread owns ONE channel vector for the whole
  // block, so its 8 scale/shift pairs live in registers and the inner loop is load - 8 FMA - 8 SiLU - store.
  __shared__ float s_mean[32];
  __shared__ float s_rstd[32];
  __shared__ float s_mean2[32];
  __shared__ float s_rstd2[32];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int cpg = C / a.groups;
  const int qpg = cpg >> 2;
  const int Q1 = a.C1 >> 2, Q2 = a.C2 >> 2;
  if (threadIdx.x < a.groups) {
    const int g = threadIdx.x;
    double s = 0.0, ss = 0.0;
    for (int i = 0; i < qpg; ++i) {
      const int q = g * qpg + i;
      const double* p = (q < Q1) ? (a.st1 + ((size_t)b * Q1 + q) * 2) : (a.st2 + ((size_t)b * Q2 + (q - Q1)) * 2);
      s += p[0];
      ss += p[1];
    }
    const double n = (double)a.HW * cpg;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  if (DUAL && threadIdx.x >= 32 && threadIdx.x < 32 + a.groups2) {
    const int g = threadIdx.x - 32;
    const int cpg2 = C / a.groups2, qpg2 = cpg2 >> 2;
    double s = 0.0, ss = 0.0;
    for (int i = 0; i < qpg2; ++i) {
      const double* p = a.st1 + ((size_t)b * Q1 + g * qpg2 + i) * 2;
      s += p[0];
      ss += p[1];
    }
    const double n = (double)a.HW * cpg2;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean2[g] = (float)mean;
    s_rstd2[g] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int nvec1 = a.C1 >> 3;
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;       // pixel slot of this thread inside one block iteration
  const int pstride = blockDim.x / nvec;     // pixels per block iteration
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vec * 8 + k;
    const int g = c / cpg;
    sc[k] = a.gamma[c] * s_rstd[g];
    sh[k] = a.beta[c] - s_mean[g] * sc[k];
  }
  float sc2[DUAL ? 8 : 1], sh2[DUAL ? 8 : 1];
  if constexpr (DUAL) {
    const int cpg2 = C / a.groups2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = vec * 8 + k;
      const int g = c / cpg2;
      sc2[k] = a.gamma2[c] * s_rstd2[g];
      sh2[k] = a.beta2[c] - s_mean2[g] * sc2[k];
    }
  }
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base_pix = (size_t)b * a.HW + pix0;
  const uint16_t* src = (vec < nvec1) ? a.x1 + base_pix * a.p1 + vec * 8 : a.x2 + base_pix * a.p2 + (vec - nvec1) * 8;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  uint16_t* dst = a.out + base_pix * a.pout + vec * 8;
  uint16_t* dst2 = DUAL ? a.out2 + base_pix * a.pout2 + vec * 8 : nullptr;

  for (int p0 = prow; p0 < npix; p0 += pstride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) v[u] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)p * sp));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        const float2 f0 = unpack2<DT>(v[u].x), f1 = unpack2<DT>(v[u].y);
        const float2 f2 = unpack2<DT>(v[u].z), f3 = unpack2<DT>(v[u].w);
        float y[8];
        y[0] = fmaf(f0.x, sc[0], sh[0]);
        y[1] = fmaf(f0.y, sc[1], sh[1]);
        y[2] = fmaf(f1.x, sc[2], sh[2]);
        y[3] = fmaf(f1.y, sc[3], sh[3]);
        y[4] = fmaf(f2.x, sc[4], sh[4]);
        y[5] = fmaf(f2.y, sc[5], sh[5]);
        y[6] = fmaf(f3.x, sc[6], sh[6]);
        y[7] = fmaf(f3.y, sc[7], sh[7]);
        if (a.silu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = silu_fast(y[k]);
        }
        uint4 o;
        o.x = pack2<DT>(y[0], y[1]);
        o.y = pack2<DT>(y[2], y[3]);
        o.z = pack2<DT>(y[4], y[5]);
        o.w = pack2<DT>(y[6], y[7]);
        *reinterpret_cast<uint4*>(dst + (size_t)p * a.pout) = o;
        if constexpr (DUAL) {
          float z[8];
          z[0] = fmaf(f0.x, sc2[0], sh2[0]);
          z[1] = fmaf(f0.y, sc2[1], sh2[1]);
          z[2] = fmaf(f1.x, sc2[2], sh2[2]);
          z[3] = fmaf(f1.y, sc2[3], sh2[3]);
          z[4] = fmaf(f2.x, sc2[4], sh2[4]);
          z[5] = fmaf(f2.y, sc2[5], sh2[5]);
          z[6] = fmaf(f3.x, sc2[6], sh2[6]);
          z[7] = fmaf(f3.y, sc2[7], sh2[7]);
          if (a.silu2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = silu_fast(z[k]);
          }
          uint4 o2;
          o2.x = pack2<DT>(z[0], z[1]);
          o2.y = pack2<DT>(z[2], z[3]);
          o2.z = pack2<DT>(z[4], z[5]);
          o2.w = pack2<DT>(z[6], z[7]);
          *reinterpret_cast<uint4*>(dst2 + (size_t)p * a.pout2) = o2;
        }
      }
    }
  }
}

// out[b, h, w, :] = in[b, h/2, w/2, :]  (F.interpolate(scale_factor=2, mode="nearest") of Upsample2D)
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint16_t* __restrict__ in,
                                                         uint16_t* __restrict__ out, int B, int H, int W, int C,
                                                         int pin, int pout) {
  pdl_prologue();
  const int nvec = C >> 3;
  const size_t total = (size_t)B * H * W * nvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    size_t p = i / nvec;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    const uint4 val = __ldcs(reinterpret_cast<const uint4*>(in + (((size_t)b * H + h) * W + w) * pin + v * 8));
    uint16_t* o = out + (((size_t)b * 2 * H + 2 * h) * (2 * W) + 2 * w) * pout + v * 8;
    *reinterpret_cast<uint4*>(o) = val;
    *reinterpret_cast<uint4*>(o + pout) = val;
    *reinterpret_cast<uint4*>(o + (size_t)2 * W * pout) = val;
    *reinterpret_cast<uint4*>(o + (size_t)2 * W * pout + pout) = val;
  }
}

}  // namespace
}  // namespace d3

namespace d3 {
namespace {
int launch_gn(GnArgs& a, int B, int dtype, bool dual, cudaStream_t stream) {
  const int C = a.C1 + a.C2;
  // at least 32 KB of work per block
  // ~8 blocks per SM over the whole batch, 8 independent 16-byte loads per thread in flight.  (Alone, the pass is
  // faster with 64 blocks per SM - 6.4 vs 5.8 TB/s - but inside the power-capped UNet forward, between two convs, 8 is
  // the better setting: 6.19 vs 6.48 ms of GroupNorm per forward.  D3_GN_BPS overrides.)
  static int bps = -1;
  if (bps < 0) { const char* e = getenv("D3_GN_BPS"); bps = e ? atoi(e) : 8; }
  int blocks_per_sample = ceil_div(num_sms() * bps, B);
  int ppb = ceil_div(a.HW, blocks_per_sample);
  const int min_ppb = ceil_div(16384, C);
  if (ppb < min_ppb) ppb = min_ppb;
  a.pix_per_block = ppb;
  dim3 grid(ceil_div(a.HW, ppb), B);
  const int nvec = C / 8;
  const int threads = nvec * (kGnThreads / nvec > 0 ? kGnThreads / nvec : 1);  // multiple of nvec, <= 256 (nvec <= 128)
  if (dual) {
    if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_apply_kernel<kF16, true>, dim3(grid), dim3(threads), 0, stream, a));
    else D3_CHECK_CUDA(launch_pdl(gn_apply_kernel<kBF16, true>, dim3(grid), dim3(threads), 0, stream, a));
  } else {
    if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_apply_kernel<kF16>, dim3(grid), dim3(threads), 0, stream, a));
    else D3_CHECK_CUDA(launch_pdl(gn_apply_kernel<kBF16>, dim3(grid), dim3(threads), 0, stream, a));
  }
  D3_LAUNCH_CHECK();
  return D3_OK;
}
}  // namespace
}  // namespace d3

extern "C" int d3_groupnorm_silu(const void* x1, int C1, int p1, const double* stats1, const void* x2, int C2, int p2,
                                 const double* stats2, const float* gamma, const float* beta, int B, int HW,
                                 int groups, float eps, int apply_silu, void* out, int pout, int dtype, void* stream) {
  using namespace d3;
  const int C = C1 + C2;
  D3_REQUIRE(x1 && stats1 && gamma && beta && out, "groupnorm: null pointer");
  D3_REQUIRE(C2 == 0 || (x2 && stats2), "groupnorm: second source missing");
  if (dtype == D3_DTYPE_F32)
    return groupnorm_f32(x1, C1, p1, stats1, x2, C2, p2, stats2, gamma, beta, B, HW, groups, eps, apply_silu, out, pout,
                         static_cast<cudaStream_t>(stream));
  D3_REQUIRE(C <= kMaxC && C % groups == 0 && (C / groups) % 4 == 0 && groups <= 32,
             "groupnorm: unsupported C=%d groups=%d", C, groups);
  D3_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && p1 % 8 == 0 && (C2 == 0 || p2 % 8 == 0) && pout % 8 == 0,
             "groupnorm: channel counts / pitches must be multiples of 8");
  GnArgs a{};
  a.x1 = static_cast<const uint16_t*>(x1);
  a.x2 = static_cast<const uint16_t*>(x2);
  a.st1 = stats1;
  a.st2 = stats2;
  a.gamma = gamma;
  a.beta = beta;
  a.out = static_cast<uint16_t*>(out);
  a.C1 = C1;
  a.C2 = C2;
  a.p1 = p1;
  a.p2 = p2;
  a.pout = pout;
  a.HW = HW;
  a.groups = groups;
  a.eps = eps;
  a.silu = apply_silu;
  return launch_gn(a, B, dtype, false, static_cast<cudaStream_t>(stream));
}

extern "C" int d3_groupnorm_silu_dual(const void* x, int C, int pitch, const double* stats, int B, int HW, float eps,
                                      const float* gamma_a, const float* beta_a, int groups_a, int silu_a, void* out_a,
                                      int pitch_a, const float* gamma_b, const float* beta_b, int groups_b, int silu_b,
                                      void* out_b, int pitch_b, int dtype, void* stream) {
  using namespace d3;
  D3_REQUIRE(x && stats && gamma_a && beta_a && out_a && gamma_b && beta_b && out_b, "groupnorm_dual: null pointer");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "groupnorm_dual: 16-bit dtypes only (got %d)", dtype);
  D3_REQUIRE(C <= kMaxC && C % 8 == 0 && pitch % 8 == 0 && pitch_a % 8 == 0 && pitch_b % 8 == 0,
             "groupnorm_dual: channel count / pitches must be multiples of 8");
  D3_REQUIRE(groups_a >= 1 && groups_a <= 32 && C % groups_a == 0 && (C / groups_a) % 4 == 0 && groups_b >= 1 &&
                 groups_b <= 32 && C % groups_b == 0 && (C / groups_b) % 4 == 0,
             "groupnorm_dual: unsupported grouping C=%d groups=%d/%d", C, groups_a, groups_b);
  GnArgs a{};
  a.x1 = static_cast<const uint16_t*>(x);
  a.st1 = stats;
  a.gamma = gamma_a;
  a.beta = beta_a;
  a.out = static_cast<uint16_t*>(out_a);
  a.C1 = C;
  a.p1 = pitch;
  a.pout = pitch_a;
  a.HW = HW;
  a.groups = groups_a;
  a.eps = eps;
  a.silu = silu_a;
  a.gamma2 = gamma_b;
  a.beta2 = beta_b;
  a.out2 = static_cast<uint16_t*>(out_b);
  a.pout2 = pitch_b;
  a.groups2 = groups_b;
  a.silu2 = silu_b;
  return launch_gn(a, B, dtype, true, static_cast<cudaStream_t>(stream));
}

extern "C" int d3_upsample_nearest2x(const void* in, int B, int H, int W, int C, int pin, void* out, int pout,
                                     void* stream) {
  using namespace d3;
  D3_REQUIRE(in && out && C % 8 == 0 && pin % 8 == 0 && pout % 8 == 0, "upsample: bad arguments");
  const size_t total = (size_t)B * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  const int cap = num_sms() * 16;
  if (blocks > cap) blocks = cap;
  D3_CHECK_CUDA(launch_pdl(upsample2x_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), B, H, W, C, pin, pout));
  D3_LAUNCH_CHECK();
  return D3_OK;
}
