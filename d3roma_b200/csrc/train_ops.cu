// Backward / optimizer kernels of the training step (reference: /root/reference/train.py:157-538 -
// `accelerator.backward(loss)` :455, `clip_grad_norm_` :518, `optimizer.step()` :524 with torch.optim.AdamW :862).
//
// Everything here is a bandwidth kernel over the same 16-bit NHWC activations the forward kernels use (the dense
// contractions of the backward pass - data gradients and weight gradients of the convolutions - run on the tcgen05
// kernels in conv_igemm.cu / conv_wgrad.cu):
//   d3_groupnorm_silu_bwd  backward of GroupNorm(+SiLU) over a (virtual) concat of two sources: per-(sample, channel)
//                          sums in one pass, dx (+ an optional addend, + optional per-(sample, channel) sums of dx for
//                          the bias / time-embedding gradients) in a second pass
//   d3_colsum_nhwc16       per-(sample, channel) sums of a tensor (bias gradients)
//   d3_add_nhwc16, d3_sumpool2x_nhwc16 (backward of the nearest-2x upsample)
//   d3_conv_out_bwd        backward of conv_norm_out + SiLU + conv_out's 128 -> dc convolution
//   d3_masked_loss_grad    the masked, per-sample weighted MSE / L1 loss of train.py:322-417 and its gradient
//   d3_linear_small_bwd    backward of the small fp32 Linear layers (time MLP, time_emb_proj)
//   d3_sumsq, d3_adamw_step, d3_gather_add, d3_pack_taps16   gradient norm / clipping, AdamW, merged biases, weight re-packing
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

D3_DEVINL float sigmoid_f(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }
// d/dz [z * sigmoid(z)]
D3_DEVINL float dsilu_f(float z) {
  const float s = sigmoid_f(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// one-MUFU forms for the bandwidth kernels (tanh.approx: ~2^-11 relative error, at the rounding of the 16-bit value
// the result is stored to): sigmoid(z) = 0.5 + 0.5 tanh(z / 2)
D3_DEVINL float sigmoid_fast(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}
D3_DEVINL float dsilu_fast(float z) {
  const float s = sigmoid_fast(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

template <int DT>
D3_DEVINL void unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = unpack2<DT>(v.x), b = unpack2<DT>(v.y), c = unpack2<DT>(v.z), d = unpack2<DT>(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <int DT>
D3_DEVINL uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack2<DT>(f[0], f[1]);
  o.y = pack2<DT>(f[2], f[3]);
  o.z = pack2<DT>(f[4], f[5]);
  o.w = pack2<DT>(f[6], f[7]);
  return o;
}

constexpr int kBwThreads = 256;
constexpr int kMaxAddends = 3;

// ----------------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) backward
// ----------------------------------------------------------------------------------------------------------------
struct GnBwdArgs {
  const uint16_t* x1;
  const uint16_t* x2;
  const double* st1;
  const double* st2;
  const float* gamma;
  const float* beta;
  const uint16_t* dy;
  const uint16_t* addend[kMaxAddends];   // optional tensors added to dx (gradients that reached the input on other paths)
  int padd[kMaxAddends];
  uint16_t* dx;
  float* dgamma;
  float* dbeta;
  float* colsum;
  long long colsum_bstride;   // elements between the rows of consecutive samples (0: all samples add into one row)
  float* ws;   // [B][C][2]: sum over pixels of dz, of dz * xhat
  int C1, C2, p1, p2, pdy, pdx;
  int colsum_C;   // leading channels whose sums are wanted
  int HW, groups;
  float eps;
  int silu;
  int pix_per_block;
};

D3_DEVINL void gn_group_moments(const GnBwdArgs& a, int b, float* s_mean, float* s_rstd) {
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
}

// pass 1: ws[b][c] += (sum_p dz, sum_p dz * xhat),  dz = dy * silu'(z),  z = xhat * gamma + beta.
// Per-thread constants are folded so that two blocks of 256 threads fit an SM without spills:
//   z = zs * x + zb (zs = rstd * gamma, zb = beta - mean * zs);  sum dz * xhat = rstd * sum dz * x - mean * rstd * sum dz.
template <int DT>
__global__ void __launch_bounds__(kBwThreads, 2) gn_bwd_reduce_kernel(const GnBwdArgs a) {
  pdl_prologue();
  __shared__ float s_mean[32];
  __shared__ float s_rstd[32];
  __shared__ float s_red[kBwThreads * 16];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int cpg = C / a.groups;
  gn_group_moments(a, b, s_mean, s_rstd);
  __syncthreads();
  const int nvec = C >> 3, nvec1 = a.C1 >> 3;
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int pstride = blockDim.x / nvec;
  float zs[8], zb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vec * 8 + k;
    const int g = c / cpg;
    zs[k] = a.gamma[c] * s_rstd[g];
    zb[k] = a.beta[c] - s_mean[g] * zs[k];
  }
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base_pix = (size_t)b * a.HW + pix0;
  const uint16_t* src = (vec < nvec1) ? a.x1 + base_pix * a.p1 + vec * 8 : a.x2 + base_pix * a.p2 + (vec - nvec1) * 8;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  const uint16_t* gy = a.dy + base_pix * a.pdy + vec * 8;
  float s1[8], s2[8];   // sum dz, sum dz * x
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  constexpr int U = 4;
  for (int p0 = prow; p0 < npix; p0 += pstride * U) {
    uint4 xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        xv[u] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)p * sp));
        gv[u] = __ldcs(reinterpret_cast<const uint4*>(gy + (size_t)p * a.pdy));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        float x[8], g[8];
        unpack8<DT>(xv[u], x);
        unpack8<DT>(gv[u], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz = g[k];
          if (a.silu) dz *= dsilu_fast(fmaf(x[k], zs[k], zb[k]));
          s1[k] += dz;
          s2[k] = fmaf(dz, x[k], s2[k]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_red[threadIdx.x * 16 + k] = s1[k];
    s_red[threadIdx.x * 16 + 8 + k] = s2[k];
  }
  __syncthreads();
  if (prow == 0) {
    for (int rr = 1; rr < pstride; ++rr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1[k] += s_red[(rr * nvec + vec) * 16 + k];
        s2[k] += s_red[(rr * nvec + vec) * 16 + 8 + k];
      }
    }
    float* w = a.ws + ((size_t)b * C + vec * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (vec * 8 + k) / cpg;
      const float r = s_rstd[g], mr = s_mean[g] * r;
      atomicAdd(w + 2 * k, s1[k]);
      atomicAdd(w + 2 * k + 1, fmaf(r, s2[k], -mr * s1[k]));   // sum dz * xhat
    }
  }
}

// pass 2: dx = rstd * (gamma * dz - mean_g(gamma * dz) - xhat * mean_g(gamma * dz * xhat)) (+ addends)
//            = zs * dz + e * x + f   with  e = kd * rstd,  f = kb - kd * mean * rstd  (kb, kd: the two group means, scaled)
// NADD: how many addend pointers the instantiation reads (0, 1 or kMaxAddends); U: pixels in flight per thread - the
// fewer streams, the deeper the pipeline the 128-register budget allows (ncu: the U = 2 form of the no-addend case sat at
// 4.1 TB/s with every warp stalled on its loads)
template <int DT, int NADD, int U>
__global__ void __launch_bounds__(kBwThreads, 2) gn_bwd_apply_kernel(const GnBwdArgs a) {
  pdl_prologue();
  __shared__ float s_mean[32];
  __shared__ float s_rstd[32];
  __shared__ float s_b[32];
  __shared__ float s_d[32];
  __shared__ float s_red[kBwThreads * 8];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int cpg = C / a.groups;
  gn_group_moments(a, b, s_mean, s_rstd);
  __syncthreads();
  if (threadIdx.x < a.groups) {
    const int g = threadIdx.x;
    float S1 = 0.f, S2 = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int c = g * cpg + i;
      const float gm = a.gamma[c];
      S1 = fmaf(gm, a.ws[((size_t)b * C + c) * 2], S1);
      S2 = fmaf(gm, a.ws[((size_t)b * C + c) * 2 + 1], S2);
    }
    const float inv_n = 1.0f / ((float)a.HW * (float)cpg);
    s_b[g] = -s_rstd[g] * S1 * inv_n;
    s_d[g] = -s_rstd[g] * S2 * inv_n;
  }
  __syncthreads();
  const int nvec = C >> 3, nvec1 = a.C1 >> 3;
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int pstride = blockDim.x / nvec;
  if (blockIdx.x == 0 && prow == 0 && a.dgamma != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = vec * 8 + k;
      atomicAdd(a.dbeta + c, a.ws[((size_t)b * C + c) * 2]);
      atomicAdd(a.dgamma + c, a.ws[((size_t)b * C + c) * 2 + 1]);
    }
  }
  float zs[8], zb[8], ce[8], cf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vec * 8 + k;
    const int g = c / cpg;
    const float r = s_rstd[g], mr = s_mean[g] * r;
    zs[k] = a.gamma[c] * r;
    zb[k] = a.beta[c] - s_mean[g] * zs[k];
    ce[k] = s_d[g] * r;
    cf[k] = s_b[g] - s_d[g] * mr;
  }
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base_pix = (size_t)b * a.HW + pix0;
  const uint16_t* src = (vec < nvec1) ? a.x1 + base_pix * a.p1 + vec * 8 : a.x2 + base_pix * a.p2 + (vec - nvec1) * 8;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  const uint16_t* gy = a.dy + base_pix * a.pdy + vec * 8;
  const uint16_t* ad[NADD > 0 ? NADD : 1];
#pragma unroll
  for (int i = 0; i < NADD; ++i) ad[i] = a.addend[i] ? a.addend[i] + base_pix * a.padd[i] + vec * 8 : nullptr;
  uint16_t* dst = a.dx + base_pix * a.pdx + vec * 8;
  const bool want_cs = a.colsum != nullptr && vec * 8 < a.colsum_C;
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
  for (int p0 = prow; p0 < npix; p0 += pstride * U) {
    uint4 xv[U], gv[U], av[U][NADD > 0 ? NADD : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        xv[u] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)p * sp));
        gv[u] = __ldcs(reinterpret_cast<const uint4*>(gy + (size_t)p * a.pdy));
#pragma unroll
        for (int i = 0; i < NADD; ++i)
          if (ad[i]) av[u][i] = __ldcs(reinterpret_cast<const uint4*>(ad[i] + (size_t)p * a.padd[i]));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        float x[8], g[8], o[8];
        unpack8<DT>(xv[u], x);
        unpack8<DT>(gv[u], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz = g[k];
          if (a.silu) dz *= dsilu_fast(fmaf(x[k], zs[k], zb[k]));
          o[k] = fmaf(zs[k], dz, fmaf(ce[k], x[k], cf[k]));
        }
#pragma unroll
        for (int i = 0; i < NADD; ++i) {
          if (ad[i]) {
            float e[8];
            unpack8<DT>(av[u][i], e);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += e[k];
          }
        }
        const uint4 pk = pack8<DT>(o);
        *reinterpret_cast<uint4*>(dst + (size_t)p * a.pdx) = pk;
        if (want_cs) {
          float rr[8];
          unpack8<DT>(pk, rr);   // sums of the values as stored
#pragma unroll
          for (int k = 0; k < 8; ++k) cs[k] += rr[k];
        }
      }
    }
  }
  if (a.colsum != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = cs[k];
    __syncthreads();
    if (prow == 0 && want_cs) {
      for (int rr = 1; rr < pstride; ++rr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] += s_red[(rr * nvec + vec) * 8 + k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(a.colsum + (size_t)b * a.colsum_bstride + vec * 8 + k, cs[k]);
    }
  }
}

// out[b][c] += sum over pixels of x[b, p, c]
template <int DT>
__global__ void __launch_bounds__(kBwThreads) colsum_kernel(const uint16_t* __restrict__ x, int pitch, int HW, int C,
                                                            int pix_per_block, float* __restrict__ out,
                                                            long long out_bstride) {
  pdl_prologue();
  __shared__ float s_red[kBwThreads * 8];
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int pstride = blockDim.x / nvec;
  const int pix0 = blockIdx.x * pix_per_block;
  int npix = HW - pix0;
  if (npix > pix_per_block) npix = pix_per_block;
  const uint16_t* src = x + ((size_t)b * HW + pix0) * pitch + vec * 8;
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
  for (int p = prow; p < npix; p += pstride) {
    float f[8];
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(src + (size_t)p * pitch)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = cs[k];
  __syncthreads();
  if (prow == 0) {
    for (int rr = 1; rr < pstride; ++rr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += s_red[(rr * nvec + vec) * 8 + k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(out + (size_t)b * out_bstride + vec * 8 + k, cs[k]);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) add16_kernel(const uint16_t* __restrict__ a, int pa, const uint16_t* __restrict__ b,
                                                    int pb, uint16_t* __restrict__ out, int po, size_t npix, int C) {
  pdl_prologue();
  const int nvec = C >> 3;
  const size_t total = npix * nvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const size_t p = i / nvec;
    float x[8], y[8];
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(a + p * pa + v * 8)), x);
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(b + p * pb + v * 8)), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] += y[k];
    *reinterpret_cast<uint4*>(out + p * po + v * 8) = pack8<DT>(x);
  }
}

// out = a + b over [B, HW, C] with the per-(sample, channel) sums of the (rounded) result added to colsum: the last add of
// a conv output's gradient and that conv's bias gradient in one pass
template <int DT>
__global__ void __launch_bounds__(kBwThreads) add_colsum_kernel(const uint16_t* __restrict__ a, int pa,
                                                                const uint16_t* __restrict__ b2, int pb,
                                                                uint16_t* __restrict__ out, int po, int HW, int C,
                                                                int pix_per_block, float* __restrict__ colsum,
                                                                long long cs_bstride) {
  pdl_prologue();
  __shared__ float s_red[kBwThreads * 8];
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int vec = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int pstride = blockDim.x / nvec;
  const int pix0 = blockIdx.x * pix_per_block;
  int npix = HW - pix0;
  if (npix > pix_per_block) npix = pix_per_block;
  const size_t base = (size_t)b * HW + pix0;
  const uint16_t* sa = a + base * pa + vec * 8;
  const uint16_t* sb = b2 + base * pb + vec * 8;
  uint16_t* so = out + base * po + vec * 8;
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
  constexpr int U = 4;
  for (int p0 = prow; p0 < npix; p0 += pstride * U) {
    uint4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        va[u] = __ldcs(reinterpret_cast<const uint4*>(sa + (size_t)p * pa));
        vb[u] = __ldcs(reinterpret_cast<const uint4*>(sb + (size_t)p * pb));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * pstride;
      if (p < npix) {
        float x[8], y[8];
        unpack8<DT>(va[u], x);
        unpack8<DT>(vb[u], y);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] += y[k];
        const uint4 pk = pack8<DT>(x);
        *reinterpret_cast<uint4*>(so + (size_t)p * po) = pk;
        unpack8<DT>(pk, y);
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] += y[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = cs[k];
  __syncthreads();
  if (prow == 0) {
    for (int rr = 1; rr < pstride; ++rr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += s_red[(rr * nvec + vec) * 8 + k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(colsum + (size_t)b * cs_bstride + vec * 8 + k, cs[k]);
  }
}

// out[b, i, j, :] = sum of the 2x2 block in[b, 2i + {0,1}, 2j + {0,1}, :] (+ addend): backward of nearest-2x upsampling
template <int DT>
__global__ void __launch_bounds__(256) sumpool2x_kernel(const uint16_t* __restrict__ in, int pin, const uint16_t* __restrict__ addend,
                                                        int padd, uint16_t* __restrict__ out, int pout, int B, int H, int W,
                                                        int C) {
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
    const uint16_t* s = in + (((size_t)b * 2 * H + 2 * h) * (2 * W) + 2 * w) * pin + v * 8;
    float acc[8], f[8];
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(s)), acc);
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(s + pin)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(s + (size_t)2 * W * pin)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
    unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(s + (size_t)2 * W * pin + pin)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
    const size_t po = ((size_t)b * H + h) * W + w;
    if (addend) {
      unpack8<DT>(__ldg(reinterpret_cast<const uint4*>(addend + po * padd + v * 8)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    *reinterpret_cast<uint4*>(out + po * pout + v * 8) = pack8<DT>(acc);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// conv_norm_out + SiLU + conv_out (128 -> dc, 3x3) backward
// ----------------------------------------------------------------------------------------------------------------
struct ConvOutBwdArgs {
  const uint16_t* h;     // pre-norm activation [B,H,W,128]
  int pitch;
  const double* stats;   // quad statistics of h
  const float* gamma;
  const float* beta;
  const float* w;        // [dc][9][128]
  const float* dy;       // fp32 NCHW [B,dc,H,W]
  uint16_t* da;          // grad wrt SiLU(GN(h)), [B,H,W,128] (written by pass 0 only)
  int pda;
  float* dw;             // [dc][9][128], +=
  int B, H, W, dc;
  float eps;
  int pass_d;            // this launch accumulates dw[pass_d]; pass 0 also writes da
  int rows_per_block;
};

template <int DT>
__global__ void __launch_bounds__(256, 2) conv_out_bwd_kernel(const ConvOutBwdArgs a) {
  pdl_prologue();
  __shared__ float s_mean[32];
  __shared__ float s_rstd[32];
  __shared__ float s_w[3 * 9 * 128];
  __shared__ float s_dw[9 * 128];
  const int b = blockIdx.y;
  if (threadIdx.x < 32) {
    const double* p = a.stats + ((size_t)b * 32 + threadIdx.x) * 2;
    const double n = (double)a.H * a.W * 4.0;
    const double mean = p[0] / n;
    double var = p[1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  // weights staged as [d * 9 + t][half][vec][4]: the 16 channel-vector lanes of a pixel read consecutive 16-byte
  // words (the straight [..][128] layout made every LDS.128 a 4-way bank conflict: ncu, 149 M conflicts per launch)
  for (int i = threadIdx.x; i < a.dc * 9 * 128; i += 256) {
    const int row = i >> 7, c = i & 127;
    s_w[row * 128 + ((c >> 2) & 1) * 64 + (c >> 3) * 4 + (c & 3)] = a.w[i];
  }
  for (int i = threadIdx.x; i < 9 * 128; i += 256) s_dw[i] = 0.f;
  __syncthreads();
  const int vec = threadIdx.x & 15;
  const int pslot = threadIdx.x >> 4;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vec * 8 + k;
    const int g = c >> 2;
    sc[k] = a.gamma[c] * s_rstd[g];
    sh[k] = a.beta[c] - s_mean[g] * sc[k];
  }
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  const int y0 = blockIdx.x * a.rows_per_block;
  const int y1 = min(a.H, y0 + a.rows_per_block);
  const int npix = (y1 - y0) * a.W;
  const size_t plane = (size_t)a.H * a.W;
  // software-pipelined over this thread's pixels: the activation vector of the next pixel is in flight while the current
  // one is processed
  int pi = pslot;
  bool have = pi < npix;
  uint4 h_next = make_uint4(0u, 0u, 0u, 0u);
  if (have) {
    const int yn = y0 + pi / a.W, xn = pi % a.W;
    h_next = __ldg(reinterpret_cast<const uint4*>(a.h + (((size_t)b * a.H + yn) * a.W + xn) * a.pitch + vec * 8));
  }
  while (have) {
    const uint4 h_cur = h_next;
    const int y = y0 + pi / a.W;
    const int x = pi % a.W;
    const size_t pix = ((size_t)b * a.H + y) * a.W + x;
    pi += 16;
    have = pi < npix;
    if (have) {
      const int yn = y0 + pi / a.W, xn = pi % a.W;
      h_next = __ldg(reinterpret_cast<const uint4*>(a.h + (((size_t)b * a.H + yn) * a.W + xn) * a.pitch + vec * 8));
    }
    float hv[8], av[8];
    unpack8<DT>(h_cur, hv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z = fmaf(hv[k], sc[k], sh[k]);
      av[k] = z * sigmoid_f(z);
    }
    float dav[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dav[k] = 0.f;
    for (int d = 0; d < a.dc; ++d) {
      if (a.pass_d != 0 && d != a.pass_d) continue;
      const float* gp = a.dy + ((size_t)b * a.dc + d) * plane;
      float g[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        // forward: m[y][x] += w[t] * a[y + kh - 1][x + kw - 1]  =>  this pixel meets dy[y - kh + 1][x - kw + 1] through tap t
        const int yy = y - (t / 3) + 1, xx = x - (t % 3) + 1;
        g[t] = (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) ? __ldg(gp + (size_t)yy * a.W + xx) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (a.pass_d == 0) {
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[(d * 9 + t) * 128 + vec * 4]);
          const float4 w1 = *reinterpret_cast<const float4*>(&s_w[(d * 9 + t) * 128 + 64 + vec * 4]);
          dav[0] = fmaf(w0.x, g[t], dav[0]); dav[1] = fmaf(w0.y, g[t], dav[1]);
          dav[2] = fmaf(w0.z, g[t], dav[2]); dav[3] = fmaf(w0.w, g[t], dav[3]);
          dav[4] = fmaf(w1.x, g[t], dav[4]); dav[5] = fmaf(w1.y, g[t], dav[5]);
          dav[6] = fmaf(w1.z, g[t], dav[6]); dav[7] = fmaf(w1.w, g[t], dav[7]);
        }
        if (d == a.pass_d) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(av[k], g[t], acc[t][k]);
        }
      }
    }
    if (a.pass_d == 0) *reinterpret_cast<uint4*>(a.da + pix * a.pda + vec * 8) = pack8<DT>(dav);
  }
  // reduce acc over the 16 pixel slots: lanes l and l ^ 16 hold the same channel vector, then the 8 warps meet in smem
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = acc[t][k] + __shfl_xor_sync(0xffffffffu, acc[t][k], 16);
      if (lane < 16) atomicAdd(&s_dw[t * 128 + lane * 8 + k], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 128; i += 256) atomicAdd(a.dw + (size_t)a.pass_d * 9 * 128 + i, s_dw[i]);
}

// ----------------------------------------------------------------------------------------------------------------
// loss: network_loss = loss_fn(pred * mask, target * mask); loss = mean(network_loss * weights[b])   (train.py:322-417)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const float* __restrict__ mask, const float* __restrict__ weights,
                                                        int dc, size_t plane, int l1, float gscale,
                                                        float* __restrict__ dy, double* __restrict__ loss_sum,
                                                        float* __restrict__ dbias) {
  // grid (chunks, dc, B)
  const int d = blockIdx.y, b = blockIdx.z;
  const float wb = weights ? weights[b] : 1.0f;
  const size_t off = ((size_t)b * dc + d) * plane;
  const float* mk = mask ? mask + (size_t)b * plane : nullptr;
  double ls = 0.0;
  float gs = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
    const float m = mk ? mk[i] : 1.0f;
    const float e = pred[off + i] * m - target[off + i] * m;
    float l, g;
    if (l1) {
      l = fabsf(e);
      g = (e > 0.f ? m : (e < 0.f ? -m : 0.f));
    } else {
      l = e * e;
      g = 2.0f * e * m;
    }
    g *= gscale * wb;
    dy[off + i] = g;
    ls += (double)l;
    gs += g;
  }
  __shared__ double s_l[256];
  __shared__ float s_g[256];
  s_l[threadIdx.x] = ls;
  s_g[threadIdx.x] = gs;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_l[threadIdx.x] += s_l[threadIdx.x + s];
      s_g[threadIdx.x] += s_g[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum + b, s_l[0]);
    if (dbias) atomicAdd(dbias + d, s_g[0]);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// small fp32 Linear backward: y = act(x) W^T (+ b), W [N][K]
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ x, int act_in,
                                                           const float* __restrict__ dy, int nb, int K, int N,
                                                           float* __restrict__ dw, float* __restrict__ db) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K);
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < nb; ++b) {
    float xv = x[(size_t)b * K + k];
    if (act_in) xv = xv * sigmoid_f(xv);
    const float g = dy[(size_t)b * N + n];
    acc = fmaf(g, xv, acc);
    accb += g;
  }
  dw[i] += acc;
  if (db != nullptr && k == 0) db[n] += accb;
}
// dx[b][k] (+)= act'(x) * sum_n dy[b][n] W[n][k]; blockIdx.y splits the n range (dx must be zero / hold the value to
// accumulate onto on entry: partial sums are added with atomics)
constexpr int kLinChunk = 256;
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ x, int act_in,
                                                           const float* __restrict__ dy, const float* __restrict__ w,
                                                           int nb, int K, int N, float* __restrict__ dx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nb * K) return;
  const int b = (int)(i / K), k = (int)(i % K);
  const int n0 = blockIdx.y * kLinChunk, n1 = min(N, n0 + kLinChunk);
  float acc = 0.f;
#pragma unroll 4
  for (int n = n0; n < n1; ++n) acc = fmaf(__ldg(dy + (size_t)b * N + n), __ldg(w + (size_t)n * K + k), acc);
  if (act_in) acc *= dsilu_f(x[i]);
  atomicAdd(dx + i, acc);
}

// ----------------------------------------------------------------------------------------------------------------
// optimizer
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += (double)v * (double)v;
  }
  __shared__ double s_s[256];
  s_s[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s_s[threadIdx.x] += s_s[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, s_s[0]);
}

struct AdamArgs {
  float* p;
  const float* g;
  float* m;
  float* v;
  uint16_t* p16;       // optional 16-bit copy of the updated parameters (first n16 elements)
  size_t n, n16;
  float lr, beta1, beta2, eps, wd, bc1, bc2;
  const double* sumsq; // optional: sum of squares of the (scaled) gradient, for clipping
  float inv_scale;     // 1 / loss scale
  float max_norm;      // <= 0: no clipping
  float* norm_out;     // optional: the un-scaled total gradient norm
  int f16;
};
__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a) {
  float gmul = a.inv_scale;
  if (a.sumsq != nullptr) {
    const float norm = (float)sqrt(*a.sumsq) * a.inv_scale;
    if (a.norm_out && blockIdx.x == 0 && threadIdx.x == 0) *a.norm_out = norm;
    if (!isfinite(norm)) return;   // overflowed / NaN gradients (fp16 loss scale too high): skip the step, like GradScaler
    if (a.max_norm > 0.f) {
      const float coef = a.max_norm / (norm + 1e-6f);   // torch.nn.utils.clip_grad_norm_
      if (coef < 1.0f) gmul *= coef;
    }
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = a.g[i] * gmul;
    float p = a.p[i];
    p *= 1.0f - a.lr * a.wd;
    const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
    a.m[i] = m;
    a.v[i] = v;
    const float denom = sqrtf(v) / sqrtf(a.bc2) + a.eps;
    p -= (a.lr / a.bc1) * (m / denom);
    a.p[i] = p;
    if (i < a.n16) {
      if (a.f16) a.p16[i] = __half_as_ushort(__float2half_rn(p));
      else a.p16[i] = __bfloat16_as_ushort(__float2bfloat16_rn(p));
    }
  }
}

__global__ void __launch_bounds__(256) cast16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, size_t n,
                                                     int f16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (f16) out[i] = __half_as_ushort(__float2half_rn(x[i]));
    else out[i] = __bfloat16_as_ushort(__float2bfloat16_rn(x[i]));
  }
}

// out[i] = src[ia[i]] + src[ib[i]]  (merged biases: conv2.bias + conv_shortcut.bias, time_emb_proj.bias + conv1.bias)
__global__ void __launch_bounds__(256) gather_add_kernel(const float* __restrict__ src, const int* __restrict__ ia,
                                                         const int* __restrict__ ib, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[ia[i]] + (ib[i] >= 0 ? src[ib[i]] : 0.f);
}

// out[ci][j][co] (16-bit) = map[j] >= 0 ? w[co][col0 + map[j] * Cin + ci] : 0   - the data-gradient ("transposed")
// weight packs: flipped taps for a 3x3 stride-1 conv, phase sub-sets for a stride-2 conv
struct PackTapsArgs {
  const float* w;
  long long w_pitch;
  int col0, Cin, Cout, Tout;
  int map[16];
  uint16_t* out;
  int f16;
};
__global__ void __launch_bounds__(256) pack_taps_kernel(const PackTapsArgs a) {
  __shared__ float tile[32][33];
  // grid: (Cout / 32, Cin / 32, Tout); 32 x 8 threads
  const int j = blockIdx.z;
  const int t = a.map[j];
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    float v = 0.f;
    if (t >= 0) v = a.w[(size_t)(co0 + r) * a.w_pitch + a.col0 + (size_t)t * a.Cin + ci0 + tx];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const float v = tile[tx][r];
    const size_t o = ((size_t)(ci0 + r) * a.Tout + j) * a.Cout + co0 + tx;
    a.out[o] = a.f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// dst[n][y][x] += mul * bilinear(src[n])(y, x)  - torch.nn.Upsample(size=(H, W), mode="bilinear") (align_corners=False):
// source index max(0, (dst + 0.5) * in / out - 0.5), neighbour clamped at the border  (pyramid_noise_like, utils.py:122-133)
__global__ void __launch_bounds__(256) bilinear_acc_kernel(const float* __restrict__ src, int64_t N, int h, int w,
                                                           float* __restrict__ dst, int H, int W, float mul) {
  const float rh = (float)h / (float)H, rw = (float)w / (float)W;
  const int64_t total = N * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int64_t r = i / W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    const float sy = fmaxf(rh * ((float)y + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < h - 1) ? 1 : 0, xp = (x0 < w - 1) ? 1 : 0;
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* p = src + (n * h + y0) * w + x0;
    const float v = (1.f - ly) * ((1.f - lx) * p[0] + lx * p[xp]) + ly * ((1.f - lx) * p[(size_t)yp * w] + lx * p[(size_t)yp * w + xp]);
    dst[i] += mul * v;
  }
}

__global__ void __launch_bounds__(256) sum_sumsq_kernel(const float* __restrict__ x, size_t n, double* __restrict__ out) {
  double s = 0.0, q = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    s += v;
    q += v * v;
  }
  __shared__ double s_s[256], s_q[256];
  s_s[threadIdx.x] = s;
  s_q[threadIdx.x] = q;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      s_s[threadIdx.x] += s_s[threadIdx.x + st];
      s_q[threadIdx.x] += s_q[threadIdx.x + st];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(out, s_s[0]);
    atomicAdd(out + 1, s_q[0]);
  }
}

// x /= std(x) (unbiased, torch.std), the moments taken from sums = (sum, sum of squares)
__global__ void __launch_bounds__(256) scale_by_rstd_kernel(float* __restrict__ x, size_t n, const double* __restrict__ sums) {
  const double mean = sums[0] / (double)n;
  const double var = (sums[1] - (double)n * mean * mean) / (double)(n - 1);
  const float sd = (float)sqrt(var);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = x[i] / sd;
}

// All transposed weight packs of a model in ONE launch: a device table of entries (offsets into the fp32 master buffer
// and into the 16-bit pack buffer) and the prefix sums of their 32 x 32 tile counts; a block finds its entry by binary
// search.  (121 separate launches of pack_taps_kernel cost 1.5 ms per training step, all of it launch latency.)
__global__ void __launch_bounds__(256) pack_taps_batched_kernel(const d3_pack_entry* __restrict__ entries,
                                                                const int* __restrict__ tile_prefix, int n_entries,
                                                                const float* __restrict__ w_base,
                                                                uint16_t* __restrict__ out_base, int f16) {
  __shared__ float tile[32][33];
  const int tid = blockIdx.x;
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {                      // last entry whose first tile is <= tid
    const int mid = (lo + hi + 1) >> 1;
    if (tile_prefix[mid] <= tid) lo = mid; else hi = mid - 1;
  }
  const d3_pack_entry e = entries[lo];
  int local = tid - tile_prefix[lo];
  const int nco = e.Cout / 32, nci = e.Cin / 32;
  const int cot = local % nco;
  local /= nco;
  const int cit = local % nci;
  const int j = local / nci;
  const int t = e.tap_map[j];
  const int co0 = cot * 32, ci0 = cit * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* w = w_base + e.w_offset;
  uint16_t* out = out_base + e.out_offset;
  for (int r = ty; r < 32; r += 8) {
    float v = 0.f;
    if (t >= 0) v = w[(size_t)(co0 + r) * e.w_pitch + e.col0 + (size_t)t * e.Cin + ci0 + tx];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const float v = tile[tx][r];
    const size_t o = ((size_t)(ci0 + r) * e.n_out_taps + j) * e.Cout + co0 + tx;
    out[o] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int grid_for(size_t total, int per_block = 256, int waves = 16) {
  size_t blocks = (total + per_block - 1) / per_block;
  const size_t cap = (size_t)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// block / grid geometry shared by the per-(sample, channel-vector) kernels
void vec_geometry(int B, int HW, int C, int* threads, dim3* grid, int* ppb) {
  const int nvec = C / 8;
  *threads = nvec * (kBwThreads / nvec > 0 ? kBwThreads / nvec : 1);
  int blocks_per_sample = ceil_div(num_sms() * 8, B);
  int p = ceil_div(HW, blocks_per_sample);
  const int min_ppb = ceil_div(16384, C);
  if (p < min_ppb) p = min_ppb;
  *ppb = p;
  *grid = dim3(ceil_div(HW, p), B);
}

}  // namespace
}  // namespace d3

extern "C" size_t d3_groupnorm_bwd_workspace_bytes(int B, int C) { return (size_t)B * C * 2 * sizeof(float); }

extern "C" int d3_groupnorm_silu_bwd(const void* x1, int C1, int p1, const double* stats1, const void* x2, int C2, int p2,
                                     const double* stats2, const float* gamma, const float* beta, int B, int HW,
                                     int groups, float eps, int apply_silu, const void* dy, int pdy,
                                     const void* const* addends, const int* addend_pitches, int n_addends, void* dx,
                                     int pdx, float* dgamma, float* dbeta, float* colsum, int64_t colsum_bstride,
                                     int colsum_channels, float* workspace, int dtype, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = C1 + C2;
  D3_REQUIRE(x1 && stats1 && gamma && beta && dy && dx && workspace, "groupnorm_bwd: null pointer");
  D3_REQUIRE(C2 == 0 || (x2 && stats2), "groupnorm_bwd: second source missing");
  D3_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "groupnorm_bwd: dgamma and dbeta go together");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "groupnorm_bwd: 16-bit dtypes only (got %d)", dtype);
  D3_REQUIRE(C <= 1024 && C % groups == 0 && (C / groups) % 4 == 0 && groups <= 32, "groupnorm_bwd: unsupported C=%d groups=%d",
             C, groups);
  D3_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && p1 % 8 == 0 && (C2 == 0 || p2 % 8 == 0) && pdy % 8 == 0 && pdx % 8 == 0,
             "groupnorm_bwd: channel counts / pitches must be multiples of 8");
  D3_REQUIRE(n_addends >= 0 && n_addends <= kMaxAddends && (n_addends == 0 || (addends && addend_pitches)),
             "groupnorm_bwd: at most %d addends", kMaxAddends);
  D3_REQUIRE(colsum_channels >= 0 && colsum_channels <= C && colsum_channels % 8 == 0, "groupnorm_bwd: bad colsum_channels");
  GnBwdArgs a{};
  a.x1 = static_cast<const uint16_t*>(x1);
  a.x2 = static_cast<const uint16_t*>(x2);
  a.st1 = stats1;
  a.st2 = stats2;
  a.gamma = gamma;
  a.beta = beta;
  a.dy = static_cast<const uint16_t*>(dy);
  for (int i = 0; i < n_addends; ++i) {
    D3_REQUIRE(addends[i] != nullptr && addend_pitches[i] % 8 == 0, "groupnorm_bwd: bad addend %d", i);
    a.addend[i] = static_cast<const uint16_t*>(addends[i]);
    a.padd[i] = addend_pitches[i];
  }
  a.colsum_C = colsum_channels > 0 ? colsum_channels : C;
  a.dx = static_cast<uint16_t*>(dx);
  a.dgamma = dgamma;
  a.dbeta = dbeta;
  a.colsum = colsum;
  a.colsum_bstride = colsum_bstride;
  a.ws = workspace;
  a.C1 = C1; a.C2 = C2; a.p1 = p1; a.p2 = p2; a.pdy = pdy; a.pdx = pdx;
  a.HW = HW;
  a.groups = groups;
  a.eps = eps;
  a.silu = apply_silu;
  int threads;
  dim3 grid;
  vec_geometry(B, HW, C, &threads, &grid, &a.pix_per_block);
  D3_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)B * C * 2 * sizeof(float), stream));
  if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_bwd_reduce_kernel<kF16>, grid, dim3(threads), 0, stream, a));
  else D3_CHECK_CUDA(launch_pdl(gn_bwd_reduce_kernel<kBF16>, grid, dim3(threads), 0, stream, a));
  D3_LAUNCH_CHECK();
#define D3_GNB_APPLY(NA, UU)                                                                                          \
  do {                                                                                                                \
    if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_bwd_apply_kernel<kF16, NA, UU>, grid, dim3(threads), 0, stream, a)); \
    else D3_CHECK_CUDA(launch_pdl(gn_bwd_apply_kernel<kBF16, NA, UU>, grid, dim3(threads), 0, stream, a));           \
  } while (0)
  if (n_addends == 0) D3_GNB_APPLY(0, 4);
  else if (n_addends == 1) D3_GNB_APPLY(1, 3);
  else D3_GNB_APPLY(kMaxAddends, 2);
#undef D3_GNB_APPLY
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_colsum_nhwc16(const void* x, int pitch, int B, int HW, int C, float* out, int64_t out_bstride, int dtype,
                                void* stream_) {
  using namespace d3;
  D3_REQUIRE(x && out && C % 8 == 0 && C <= 2048 && pitch % 8 == 0, "colsum: bad arguments (C=%d)", C);
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "colsum: 16-bit dtypes only");
  int threads, ppb;
  dim3 grid;
  vec_geometry(B, HW, C, &threads, &grid, &ppb);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(colsum_kernel<kF16>, grid, dim3(threads), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(x), pitch, HW, C, ppb, out, (long long)out_bstride));
  else
    D3_CHECK_CUDA(launch_pdl(colsum_kernel<kBF16>, grid, dim3(threads), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(x), pitch, HW, C, ppb, out, (long long)out_bstride));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_add_nhwc16(const void* a, int pa, const void* b, int pb, void* out, int po, int64_t npix, int C,
                             int dtype, void* stream_) {
  using namespace d3;
  D3_REQUIRE(a && b && out && C % 8 == 0 && pa % 8 == 0 && pb % 8 == 0 && po % 8 == 0, "add: bad arguments");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "add: 16-bit dtypes only");
  const int g = grid_for((size_t)npix * (C / 8));
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(add16_kernel<kF16>, dim3(g), dim3(256), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(a), pa, static_cast<const uint16_t*>(b), pb, static_cast<uint16_t*>(out), po, (size_t)npix, C));
  else
    D3_CHECK_CUDA(launch_pdl(add16_kernel<kBF16>, dim3(g), dim3(256), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(a), pa, static_cast<const uint16_t*>(b), pb, static_cast<uint16_t*>(out), po, (size_t)npix, C));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_add_colsum_nhwc16(const void* a, int pa, const void* b, int pb, void* out, int po, int B, int HW, int C,
                                    float* colsum, int64_t colsum_bstride, int dtype, void* stream_) {
  using namespace d3;
  D3_REQUIRE(a && b && out && colsum && C % 8 == 0 && C <= 2048 && pa % 8 == 0 && pb % 8 == 0 && po % 8 == 0,
             "add_colsum: bad arguments");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "add_colsum: 16-bit dtypes only");
  int threads, ppb;
  dim3 grid;
  vec_geometry(B, HW, C, &threads, &grid, &ppb);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(add_colsum_kernel<kF16>, grid, dim3(threads), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(a), pa, static_cast<const uint16_t*>(b), pb, static_cast<uint16_t*>(out), po, HW, C, ppb, colsum, (long long)colsum_bstride));
  else
    D3_CHECK_CUDA(launch_pdl(add_colsum_kernel<kBF16>, grid, dim3(threads), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(a), pa, static_cast<const uint16_t*>(b), pb, static_cast<uint16_t*>(out), po, HW, C, ppb, colsum, (long long)colsum_bstride));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_sumpool2x_nhwc16(const void* in, int pin, const void* addend, int padd, void* out, int pout, int B,
                                   int H, int W, int C, int dtype, void* stream_) {
  using namespace d3;
  D3_REQUIRE(in && out && C % 8 == 0 && pin % 8 == 0 && pout % 8 == 0 && (addend == nullptr || padd % 8 == 0),
             "sumpool2x: bad arguments");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "sumpool2x: 16-bit dtypes only");
  const int g = grid_for((size_t)B * H * W * (C / 8));
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(sumpool2x_kernel<kF16>, dim3(g), dim3(256), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(in), pin, static_cast<const uint16_t*>(addend), padd, static_cast<uint16_t*>(out), pout, B, H, W, C));
  else
    D3_CHECK_CUDA(launch_pdl(sumpool2x_kernel<kBF16>, dim3(g), dim3(256), 0, static_cast<cudaStream_t>(stream_), static_cast<const uint16_t*>(in), pin, static_cast<const uint16_t*>(addend), padd, static_cast<uint16_t*>(out), pout, B, H, W, C));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_conv_out_bwd(const void* h, int pitch, const double* stats, const float* gamma, const float* beta,
                               const float* w, const float* dy, int B, int H, int W, int dc, float eps, void* da, int pda,
                               float* dw, int dtype, void* stream_) {
  using namespace d3;
  D3_REQUIRE(h && stats && gamma && beta && w && dy && da && dw, "conv_out_bwd: null pointer");
  D3_REQUIRE(dc >= 1 && dc <= 3 && pitch % 8 == 0 && pda % 8 == 0, "conv_out_bwd: bad arguments (dc=%d)", dc);
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "conv_out_bwd: 16-bit dtypes only");
  ConvOutBwdArgs a{};
  a.h = static_cast<const uint16_t*>(h);
  a.pitch = pitch;
  a.stats = stats;
  a.gamma = gamma;
  a.beta = beta;
  a.w = w;
  a.dy = dy;
  a.da = static_cast<uint16_t*>(da);
  a.pda = pda;
  a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.dc = dc;
  a.eps = eps;
  int blocks_per_sample = ceil_div(num_sms() * 8, B);
  a.rows_per_block = ceil_div(H, blocks_per_sample);
  if (a.rows_per_block < 1) a.rows_per_block = 1;
  dim3 grid(ceil_div(H, a.rows_per_block), B);
  for (int d = 0; d < dc; ++d) {
    a.pass_d = d;
    if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(conv_out_bwd_kernel<kF16>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream_), a));
    else D3_CHECK_CUDA(launch_pdl(conv_out_bwd_kernel<kBF16>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream_), a));
    D3_LAUNCH_CHECK();
  }
  return D3_OK;
}

extern "C" int d3_masked_loss_grad(const float* pred, const float* target, const float* mask, const float* weights, int B,
                                   int dc, int64_t plane, int l1, float grad_scale, float* dy, double* loss_sum,
                                   float* dbias, void* stream_) {
  using namespace d3;
  D3_REQUIRE(pred && target && dy && loss_sum && B > 0 && dc > 0 && plane > 0, "loss_grad: bad arguments");
  int gx = (int)((plane + 256 * 8 - 1) / (256 * 8));
  const int cap = ceil_div(num_sms() * 8, B * dc);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  loss_grad_kernel<<<dim3(gx, dc, B), 256, 0, static_cast<cudaStream_t>(stream_)>>>(pred, target, mask, weights, dc,
                                                                                  (size_t)plane, l1, grad_scale, dy,
                                                                                  loss_sum, dbias);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_linear_small_bwd(const float* x, int n_rows, int K, int act_in, const float* w, int rows,
                                   const float* dy, float* dw, float* db, float* dx, int accumulate_dx, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(x && dy && n_rows > 0 && K > 0 && rows > 0, "linear_small_bwd: bad arguments");
  D3_REQUIRE(act_in == 0 || act_in == 1, "linear_small_bwd: only identity / SiLU inputs have a backward (got %d)", act_in);
  if (dw != nullptr) {
    linear_bwd_w_kernel<<<(unsigned)(((size_t)rows * K + 255) / 256), 256, 0, stream>>>(x, act_in, dy, n_rows, K, rows, dw, db);
    D3_LAUNCH_CHECK();
  }
  if (dx != nullptr) {
    D3_REQUIRE(w != nullptr, "linear_small_bwd: dx needs the weight");
    if (!accumulate_dx) D3_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_rows * K * sizeof(float), stream));
    linear_bwd_x_kernel<<<dim3((unsigned)(((size_t)n_rows * K + 255) / 256), (unsigned)ceil_div(rows, kLinChunk)), 256, 0,
                          stream>>>(x, act_in, dy, w, n_rows, K, rows, dx);
    D3_LAUNCH_CHECK();
  }
  return D3_OK;
}

extern "C" int d3_sumsq(const float* g, int64_t n, double* out, void* stream_) {
  using namespace d3;
  D3_REQUIRE(g && out && n > 0, "sumsq: bad arguments");
  sumsq_kernel<<<grid_for((size_t)n, 256 * 8, 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(g, (size_t)n, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, void* p16, int64_t n16, int dtype16,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             const double* sumsq, float inv_loss_scale, float max_norm, float* norm_out, void* stream_) {
  using namespace d3;
  D3_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw: bad arguments");
  D3_REQUIRE(n16 == 0 || (p16 && n16 <= n && (dtype16 == D3_DTYPE_BF16 || dtype16 == D3_DTYPE_F16)), "adamw: bad 16-bit copy");
  AdamArgs a{};
  a.p = p; a.g = g; a.m = m; a.v = v;
  a.p16 = static_cast<uint16_t*>(p16);
  a.n = (size_t)n;
  a.n16 = (size_t)n16;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  a.sumsq = sumsq;
  a.inv_scale = inv_loss_scale;
  a.max_norm = max_norm;
  a.norm_out = norm_out;
  a.f16 = dtype16 == D3_DTYPE_F16;
  adamw_kernel<<<grid_for((size_t)n, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_cast_f32_to_16(const float* x, void* out, int64_t n, int dtype16, void* stream_) {
  using namespace d3;
  D3_REQUIRE(x && out && n > 0 && (dtype16 == D3_DTYPE_BF16 || dtype16 == D3_DTYPE_F16), "cast16: bad arguments");
  cast16_kernel<<<grid_for((size_t)n, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, static_cast<uint16_t*>(out), (size_t)n, dtype16 == D3_DTYPE_F16);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_gather_add(const float* src, const int32_t* ia, const int32_t* ib, float* out, int n, void* stream_) {
  using namespace d3;
  D3_REQUIRE(src && ia && ib && out && n > 0, "gather_add: bad arguments");
  gather_add_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, ia, ib, out, n);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_pack_taps16(const float* w, int64_t w_pitch, int col0, int Cin, int Cout, const int32_t* tap_map,
                              int n_out_taps, void* out, int dtype16, void* stream_) {
  using namespace d3;
  D3_REQUIRE(w && tap_map && out && Cin % 32 == 0 && Cout % 32 == 0 && n_out_taps >= 1 && n_out_taps <= 16,
             "pack_taps16: bad arguments (Cin=%d Cout=%d taps=%d)", Cin, Cout, n_out_taps);
  D3_REQUIRE(dtype16 == D3_DTYPE_BF16 || dtype16 == D3_DTYPE_F16, "pack_taps16: 16-bit dtypes only");
  PackTapsArgs a{};
  a.w = w;
  a.w_pitch = w_pitch;
  a.col0 = col0;
  a.Cin = Cin;
  a.Cout = Cout;
  a.Tout = n_out_taps;
  for (int j = 0; j < n_out_taps; ++j) a.map[j] = tap_map[j];
  a.out = static_cast<uint16_t*>(out);
  a.f16 = dtype16 == D3_DTYPE_F16;
  pack_taps_kernel<<<dim3(Cout / 32, Cin / 32, n_out_taps), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_bilinear_upsample_acc(const float* src, int64_t planes, int h, int w, float* dst, int H, int W, float mul,
                                        void* stream_) {
  using namespace d3;
  D3_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bilinear_upsample_acc: bad arguments");
  bilinear_acc_kernel<<<grid_for((size_t)planes * H * W), 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, planes, h, w, dst,
                                                                                                      H, W, mul);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_div_by_std(float* x, int64_t n, double* workspace2, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(x && workspace2 && n > 1, "div_by_std: bad arguments");
  D3_CHECK_CUDA(cudaMemsetAsync(workspace2, 0, 2 * sizeof(double), stream));
  sum_sumsq_kernel<<<grid_for((size_t)n, 256 * 8, 8), 256, 0, stream>>>(x, (size_t)n, workspace2);
  D3_LAUNCH_CHECK();
  scale_by_rstd_kernel<<<grid_for((size_t)n, 256 * 4, 8), 256, 0, stream>>>(x, (size_t)n, workspace2);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_pack_taps16_batched(const d3_pack_entry* entries_dev, const int32_t* tile_prefix_dev, int n_entries,
                                      int total_tiles, const float* w_base, void* out_base, int dtype16, void* stream_) {
  using namespace d3;
  D3_REQUIRE(entries_dev && tile_prefix_dev && w_base && out_base && n_entries > 0 && total_tiles > 0,
             "pack_taps16_batched: bad arguments");
  D3_REQUIRE(dtype16 == D3_DTYPE_BF16 || dtype16 == D3_DTYPE_F16, "pack_taps16_batched: 16-bit dtypes only");
  pack_taps_batched_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      entries_dev, tile_prefix_dev, n_entries, w_base, static_cast<uint16_t*>(out_base), dtype16 == D3_DTYPE_F16);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
