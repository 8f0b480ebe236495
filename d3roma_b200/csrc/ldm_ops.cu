// Kernels of the LATENT family that the pixel-space UNet does not need: the transformer blocks of
// diffusers' UNet2DConditionModel in its SD-2.1 configuration (loaded at /root/reference/train.py:743-772, called at
// core/custom_pipelines.py:386) and the glue around them.
//
//   d3_gn_stats / d3_gn_apply   GroupNorm for ANY group width (320/32 = 10, 1920/32 = 60 ... channels per group are
//                               not sums of the conv epilogue's 4-channel quads): per-channel fp64 moments in one pass,
//                               then the normalise(+SiLU) pass over up to two concatenated sources.
//   d3_layernorm                LayerNorm over the channel dimension of [tokens][C] (BasicTransformerBlock.norm1-3).
//   d3_geglu                    GEGLU of the feed-forward: out = h * gelu(gate), [tokens][2D] -> [tokens][D].
//   d3_attention_hd64           softmax(q k^T / 8) v with head_dim 64 (flash-attention style, mma.sync m16n8k16 fed by
//                               ldmatrix from swizzled cp.async tiles); q / k / v are channel slices of the fused QKV
//                               projection the tcgen05 igemm kernel produces.
//   d3_cross_attention_small    attention against a context of <= 8 tokens (the [B, 2, 1024] empty-prompt embedding,
//                               custom_pipelines.py:241-253, 299-304): keys / values are precomputed once per loop.
//   d3_linear_small             y = act_out(W act_in(x) + b) for a handful of rows (time embedding MLP, the 22
//                               time_emb_proj layers at once, the context K / V projections).
//   d3_timestep_sincos          diffusers Timesteps(dim, flip_sin_to_cos, freq_shift).
//   d3_nchw_to_nhwc16 / d3_nhwc16_to_nchw / d3_resize_nearest_nhwc16   layout glue.
//
// All linear layers (proj_in, QKV, to_out, GEGLU proj, FF out, proj_out) run on the tcgen05 implicit-GEMM kernel as
// 1x1 convolutions over the token grid (conv_igemm.cu) with the residual add in its epilogue.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// GroupNorm, generic group width.  Thread layout of both passes: blockDim.x = nv * k where nv = (channels of this
// block's channel chunk) / 8; every thread owns one 16-byte channel vector and walks over pixels.
// ---------------------------------------------------------------------------------------------------------------------
struct GnGenArgs {
  const uint16_t* x1;
  const uint16_t* x2;
  int C1, C2, p1, p2;
  int HW;
  int nv;            // channel vectors per chunk (C / 8 / gridDim.z)
  int pix_per_block;
  double* chan;      // [B][C][2] per-channel sum / sum of squares (zero on entry for the stats pass)
  // apply pass
  const float* gamma;
  const float* beta;
  uint16_t* out;
  int pout, groups, silu;
  float eps;
};

template <int DT>
__global__ void __launch_bounds__(256) gn_stats_kernel(const GnGenArgs a) {
  pdl_prologue();
  __shared__ float s_red[256 * 16];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int vec = blockIdx.z * a.nv + threadIdx.x % a.nv;   // global channel vector
  const int prow = threadIdx.x / a.nv, pstride = blockDim.x / a.nv;
  const int nvec1 = a.C1 >> 3;
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base_pix = (size_t)b * a.HW + pix0;
  const uint16_t* src = (vec < nvec1) ? a.x1 + base_pix * a.p1 + vec * 8 : a.x2 + base_pix * a.p2 + (vec - nvec1) * 8;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  for (int p = prow; p < npix; p += pstride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)p * sp));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      s[2 * k] += f.x;
      q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
      s[2 * k + 1] += f.y;
      q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_red[threadIdx.x * 16 + k] = s[k];
    s_red[threadIdx.x * 16 + 8 + k] = q[k];
  }
  __syncthreads();
  if (prow == 0) {   // fixed-order combine of the pixel slots, then one fp64 atomic per value and block
    const int v0 = threadIdx.x;   // == threadIdx.x % nv
    double* dst = a.chan + ((size_t)b * C + (size_t)vec * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double ts = 0.0, tq = 0.0;
      for (int r = 0; r < pstride; ++r) {
        ts += (double)s_red[(r * a.nv + v0) * 16 + k];
        tq += (double)s_red[(r * a.nv + v0) * 16 + 8 + k];
      }
      atomicAdd(dst + 2 * k, ts);
      atomicAdd(dst + 2 * k + 1, tq);
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) gn_apply_gen_kernel(const GnGenArgs a) {
  pdl_prologue();
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int cpg = C / a.groups;
  {  // group moments from the per-channel moments: one warp per group, lanes over the group's channels
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int g = warp; g < a.groups && warp < nw; g += nw) {   // a trailing partial warp takes no group
      double s = 0.0, ss = 0.0;
      for (int c = lane; c < cpg; c += 32) {
        const double* p = a.chan + ((size_t)b * C + (size_t)g * cpg + c) * 2;
        s += p[0];
        ss += p[1];
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if (lane == 0) {
        const double n = (double)a.HW * cpg;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = (float)mean;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)a.eps));
      }
    }
  }
  __syncthreads();
  const int vec = blockIdx.z * a.nv + threadIdx.x % a.nv;
  const int prow = threadIdx.x / a.nv, pstride = blockDim.x / a.nv;
  const int nvec1 = a.C1 >> 3;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = vec * 8 + k;
    const int g = c / cpg;
    sc[k] = a.gamma[c] * s_rstd[g];
    sh[k] = a.beta[c] - s_mean[g] * sc[k];
  }
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base_pix = (size_t)b * a.HW + pix0;
  const uint16_t* src = (vec < nvec1) ? a.x1 + base_pix * a.p1 + vec * 8 : a.x2 + base_pix * a.p2 + (vec - nvec1) * 8;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  uint16_t* dst = a.out + base_pix * a.pout + vec * 8;
  for (int p = prow; p < npix; p += pstride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)p * sp));
    const float2 f0 = unpack2<DT>(v.x), f1 = unpack2<DT>(v.y), f2 = unpack2<DT>(v.z), f3 = unpack2<DT>(v.w);
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
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// GroupNorm in ONE kernel for the small tensors of the latent UNet (<= 40 MB: they sit in the 126 MB L2 between the two
// passes): a block owns one sample and a set of whole groups whose channels form 16-byte vectors, reduces their moments
// itself (no atomics, no statistics arena, no second launch) and normalises from the L2-resident data.
// ---------------------------------------------------------------------------------------------------------------------
struct GnOneArgs {
  const uint16_t* x1;
  const uint16_t* x2;
  int C1, C2, p1, p2, HW;
  int cpg;           // channels per group
  int gpb;           // groups per block: gpb * cpg is a multiple of 8 (whole 16-byte vectors)
  const float* gamma;
  const float* beta;
  uint16_t* out;
  int pout, silu;
  float eps;
};

template <int DT>
__global__ void __launch_bounds__(512) gn_one_kernel(const GnOneArgs a) {
  pdl_prologue();
  constexpr int kMaxG = 8;                      // groups per block
  __shared__ double s_sum[16][kMaxG][2];        // [warp][group][sum, sum of squares]
  __shared__ float s_mean[kMaxG], s_rstd[kMaxG];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * a.gpb * a.cpg;    // first channel of this block
  const int nv = a.gpb * a.cpg / 8;             // 16-byte vectors per pixel in this block's slice
  const int total = a.HW * nv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  auto src_of = [&](int pix, int vi) -> const uint16_t* {
    const int c = c0 + vi * 8;
    return c < a.C1 ? a.x1 + ((size_t)b * a.HW + pix) * a.p1 + c : a.x2 + ((size_t)b * a.HW + pix) * a.p2 + (c - a.C1);
  };
  // ---- pass 1: per-thread moments of the channels of one vector position (vi is fixed per thread when nv | blockDim)
  float s[kMaxG], q[kMaxG];
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) s[g] = q[g] = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int pix = i / nv, vi = i - pix * nv;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src_of(pix, vi)));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      const int ch = vi * 8 + 2 * k;            // channel inside the slice (cpg is even: both halves in one group)
      const int g = ch / a.cpg;
#pragma unroll
      for (int gg = 0; gg < kMaxG; ++gg)
        if (gg == g) {
          s[gg] += f.x + f.y;
          q[gg] = fmaf(f.x, f.x, fmaf(f.y, f.y, q[gg]));
        }
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    if (g < a.gpb) {
      double ds = (double)s[g], dq = (double)q[g];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, off);
        dq += __shfl_xor_sync(0xffffffffu, dq, off);
      }
      if (lane == 0) {
        s_sum[warp][g][0] = ds;
        s_sum[warp][g][1] = dq;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < a.gpb) {
    double ds = 0.0, dq = 0.0;
    for (int w = 0; w < nw; ++w) {
      ds += s_sum[w][threadIdx.x][0];
      dq += s_sum[w][threadIdx.x][1];
    }
    const double n = (double)a.HW * a.cpg;
    const double mean = ds / n;
    double var = dq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  __syncthreads();
  // ---- pass 2: normalise (+ SiLU) from the L2-resident slice
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int pix = i / nv, vi = i - pix * nv;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src_of(pix, vi)));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      const int ch = vi * 8 + 2 * k;
      const int g = ch / a.cpg;
      const int c = c0 + ch;
      const float r = s_rstd[g], m = s_mean[g];
      const float g0 = __ldg(a.gamma + c) * r, g1 = __ldg(a.gamma + c + 1) * r;
      float y0 = fmaf(f.x - m, g0, __ldg(a.beta + c)), y1 = fmaf(f.y - m, g1, __ldg(a.beta + c + 1));
      if (a.silu) {
        y0 = silu_fast(y0);
        y1 = silu_fast(y1);
      }
      o[k] = pack2<DT>(y0, y1);
    }
    *reinterpret_cast<uint4*>(a.out + ((size_t)b * a.HW + pix) * a.pout + c0 + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// smallest number of channel chunks such that a chunk has <= 128 channel vectors and the chunks are equal
int gn_chunks(int nvec) {
  for (int n = 1; n <= nvec; ++n)
    if (nvec % n == 0 && nvec / n <= 128) return n;
  return nvec;
}

int gn_gen_setup(GnGenArgs& a, const void* x1, int C1, int p1, const void* x2, int C2, int p2, int B, int HW,
                 double* chan, dim3* grid, int* threads) {
  const int C = C1 + C2;
  D3_REQUIRE(x1 && chan && C1 > 0 && C1 % 8 == 0 && C2 % 8 == 0 && p1 % 8 == 0 && (C2 == 0 || (x2 && p2 % 8 == 0)),
             "gn: channel counts / pitches must be multiples of 8");
  a.x1 = static_cast<const uint16_t*>(x1);
  a.x2 = static_cast<const uint16_t*>(x2);
  a.C1 = C1;
  a.C2 = C2;
  a.p1 = p1;
  a.p2 = p2;
  a.HW = HW;
  a.chan = chan;
  const int nvec = C / 8;
  const int chunks = gn_chunks(nvec);
  a.nv = nvec / chunks;
  const int k = 256 / a.nv > 0 ? 256 / a.nv : 1;
  *threads = a.nv * k;
  int blocks_per_sample = ceil_div(num_sms() * 4, B * chunks);
  if (blocks_per_sample < 1) blocks_per_sample = 1;
  int ppb = ceil_div(HW, blocks_per_sample);
  const int min_ppb = k * 4;    // at least 4 pixels per thread
  if (ppb < min_ppb) ppb = min_ppb;
  a.pix_per_block = ppb;
  *grid = dim3(ceil_div(HW, ppb), B, chunks);
  return D3_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm over C (<= 1280) of [M][C] rows: one warp per row, the row lives in registers (<= 5 vectors per lane).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLnMaxV = 5;

template <int DT>
__global__ void __launch_bounds__(256) layernorm_kernel(const uint16_t* __restrict__ x, int xp, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, uint16_t* __restrict__ out, int op,
                                                        long long M, int C, float eps) {
  pdl_prologue();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C >> 3;
  const uint16_t* src = x + row * xp;
  float v[kLnMaxV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + vi * 8));
      const float2 f0 = unpack2<DT>(u.x), f1 = unpack2<DT>(u.y), f2 = unpack2<DT>(u.z), f3 = unpack2<DT>(u.w);
      v[i][0] = f0.x; v[i][1] = f0.y; v[i][2] = f1.x; v[i][3] = f1.y;
      v[i][4] = f2.x; v[i][5] = f2.y; v[i][6] = f3.x; v[i][7] = f3.y;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[i][k] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q / (float)C + eps);
  uint16_t* dst = out + row * op;
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      float y[8];
      y[0] = fmaf((v[i][0] - mean) * rstd, g0.x, b0.x);
      y[1] = fmaf((v[i][1] - mean) * rstd, g0.y, b0.y);
      y[2] = fmaf((v[i][2] - mean) * rstd, g0.z, b0.z);
      y[3] = fmaf((v[i][3] - mean) * rstd, g0.w, b0.w);
      y[4] = fmaf((v[i][4] - mean) * rstd, g1.x, b1.x);
      y[5] = fmaf((v[i][5] - mean) * rstd, g1.y, b1.y);
      y[6] = fmaf((v[i][6] - mean) * rstd, g1.z, b1.z);
      y[7] = fmaf((v[i][7] - mean) * rstd, g1.w, b1.w);
      uint4 o;
      o.x = pack2<DT>(y[0], y[1]);
      o.y = pack2<DT>(y[2], y[3]);
      o.z = pack2<DT>(y[4], y[5]);
      o.w = pack2<DT>(y[6], y[7]);
      *reinterpret_cast<uint4*>(dst + vi * 8) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// GEGLU: out[m][d] = x[m][d] * gelu(x[m][D + d])   (diffusers GEGLU: proj(x).chunk(2); exact erf GELU)
// ---------------------------------------------------------------------------------------------------------------------
D3_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }

template <int DT>
__global__ void __launch_bounds__(256) geglu_kernel(const uint16_t* __restrict__ x, int xp, uint16_t* __restrict__ out,
                                                    int op, long long M, int D) {
  pdl_prologue();
  const int nvec = D >> 3;
  const long long total = M * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nvec;
    const int vi = (int)(i - m * nvec);
    const uint4 h = __ldcs(reinterpret_cast<const uint4*>(x + m * xp + vi * 8));
    const uint4 g = __ldcs(reinterpret_cast<const uint4*>(x + m * xp + D + vi * 8));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 hf = unpack2<DT>(hw[k]), gf = unpack2<DT>(gw[k]);
      o[k] = pack2<DT>(hf.x * gelu_erf(gf.x), hf.y * gelu_erf(gf.y));
    }
    *reinterpret_cast<uint4*>(out + m * op + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Self-attention, head_dim 64.  One block = 64 queries of one (sample, head), 4 warps x 16 queries; keys / values stream
// through double-buffered 64-key tiles ([key][64 dims], 128 B rows, 16-byte chunks XOR-swizzled by key % 8) filled with
// cp.async; S = Q K^T and O += P V on mma.sync.m16n8k16 with ldmatrix (K) / ldmatrix.trans (V) operand fetches; online
// softmax in the exp2 domain.
// ---------------------------------------------------------------------------------------------------------------------
D3_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int DT>
D3_DEVINL void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (DT == kBF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}
D3_DEVINL void ldsm_x4(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
D3_DEVINL void ldsm_x4_t(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
D3_DEVINL void cp_async16(uint32_t saddr, const void* gptr, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(sz) : "memory");
}
D3_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
D3_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kA64Tile = 64;                 // keys per tile
constexpr int kA64TileBytes = kA64Tile * 128;  // one K or V tile

template <int DT>
__global__ void __launch_bounds__(128) attention_hd64_kernel(const uint16_t* __restrict__ q, int qp,
                                                             const uint16_t* __restrict__ k, int kp,
                                                             const uint16_t* __restrict__ v, int vp, int N,
                                                             uint16_t* __restrict__ out, int op, float scale_log2) {
  pdl_prologue();
  __shared__ __align__(128) uint8_t s_kv[2][2][kA64TileBytes];   // [buffer][K | V]
  const int head = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint16_t* kbase = k + (size_t)b * N * kp + head * 64;
  const uint16_t* vbase = v + (size_t)b * N * vp + head * 64;
  const int ntiles = (N + kA64Tile - 1) / kA64Tile;

  auto load_tile = [&](int tile, int buf) {
    const int key0 = tile * kA64Tile;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 128 * i;      // 512 16-byte chunks per tile
      const int row = idx >> 3, ch = idx & 7;
      const bool valid = key0 + row < N;
      const size_t key = valid ? (size_t)(key0 + row) : 0;
      const uint32_t off = (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
      cp_async16(smem_u32(&s_kv[buf][0][0]) + off, kbase + key * kp + ch * 8, valid);
      cp_async16(smem_u32(&s_kv[buf][1][0]) + off, vbase + key * vp + ch * 8, valid);
    }
    cp_async_commit();
  };
  load_tile(0, 0);

  // Q fragments: 16 queries x 64 dims = 4 k-steps of (a0..a3)
  const int q0 = blockIdx.x * 64 + warp * 16;
  uint32_t qa[4][4];
  {
    const uint16_t* qb = q + (size_t)b * N * qp + head * 64;
    const bool v0 = q0 + g < N, v1 = q0 + g + 8 < N;
    const uint16_t* r0 = qb + (size_t)(v0 ? q0 + g : 0) * qp;
    const uint16_t* r1 = qb + (size_t)(v1 ? q0 + g + 8 : 0) * qp;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = v0 ? __ldg(reinterpret_cast<const uint32_t*>(r0 + 16 * kk + 2 * t)) : 0u;
      qa[kk][1] = v1 ? __ldg(reinterpret_cast<const uint32_t*>(r1 + 16 * kk + 2 * t)) : 0u;
      qa[kk][2] = v0 ? __ldg(reinterpret_cast<const uint32_t*>(r0 + 16 * kk + 8 + 2 * t)) : 0u;
      qa[kk][3] = v1 ? __ldg(reinterpret_cast<const uint32_t*>(r1 + 16 * kk + 8 + 2 * t)) : 0u;
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  const int lm = lane >> 3, lr = lane & 7;   // ldmatrix: this lane addresses row lr of matrix lm

  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) {
      load_tile(tile + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sK = smem_u32(&s_kv[buf][0][0]), sV = smem_u32(&s_kv[buf][1][0]);
    // ---- S = Q K^T : 8 key groups x 4 k-steps
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // matrices: key rows 8j..8j+7, dims 32h + 8m .. +7  ->  (b0, b1) of k-step 2h and (b0, b1) of k-step 2h+1
        uint32_t kb[4];
        ldsm_x4(sK + (uint32_t)(8 * j + lr) * 128u + (uint32_t)(((4 * h + lm) ^ lr) << 4), kb);
        mma16816<DT>(s[j], qa[2 * h][0], qa[2 * h][1], qa[2 * h][2], qa[2 * h][3], kb[0], kb[1]);
        mma16816<DT>(s[j], qa[2 * h + 1][0], qa[2 * h + 1][1], qa[2 * h + 1][2], qa[2 * h + 1][3], kb[2], kb[3]);
      }
    }
    if (tile == ntiles - 1 && (N & (kA64Tile - 1)) != 0) {   // keys >= N of the last, partial tile
      const int key0 = tile * kA64Tile;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = key0 + 8 * j + 2 * t;
        if (key >= N) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= N) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ex2_approx((m0 - mx0) * scale_log2), c1 = ex2_approx((m1 - mx1) * scale_log2);
    m0 = mx0;
    m1 = mx1;
    l0 *= c0;
    l1 *= c1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1;
    }
    const float nm0 = -m0 * scale_log2, nm1 = -m1 * scale_log2;
    // ---- O += P V : 4 steps of 16 keys x 8 dim groups
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float p[8];
      p[0] = ex2_approx(fmaf(s[2 * i][0], scale_log2, nm0));
      p[1] = ex2_approx(fmaf(s[2 * i][1], scale_log2, nm0));
      p[2] = ex2_approx(fmaf(s[2 * i][2], scale_log2, nm1));
      p[3] = ex2_approx(fmaf(s[2 * i][3], scale_log2, nm1));
      p[4] = ex2_approx(fmaf(s[2 * i + 1][0], scale_log2, nm0));
      p[5] = ex2_approx(fmaf(s[2 * i + 1][1], scale_log2, nm0));
      p[6] = ex2_approx(fmaf(s[2 * i + 1][2], scale_log2, nm1));
      p[7] = ex2_approx(fmaf(s[2 * i + 1][3], scale_log2, nm1));
      l0 += (p[0] + p[1]) + (p[4] + p[5]);
      l1 += (p[2] + p[3]) + (p[6] + p[7]);
      const uint32_t a0 = pack2<DT>(p[0], p[1]), a1 = pack2<DT>(p[2], p[3]);
      const uint32_t a2 = pack2<DT>(p[4], p[5]), a3 = pack2<DT>(p[6], p[7]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        // transposed matrices: keys 16i + 8(m & 1) .. +7, dims 8(2 n2 + (m >> 1)) .. +7 -> (b0, b1) of dim groups 2 n2, 2 n2 + 1
        uint32_t vb[4];
        const int row = 16 * i + 8 * (lm & 1) + lr;
        ldsm_x4_t(sV + (uint32_t)row * 128u + (uint32_t)(((2 * n2 + (lm >> 1)) ^ lr) << 4), vb);
        mma16816<DT>(o[2 * n2], a0, a1, a2, a3, vb[0], vb[1]);
        mma16816<DT>(o[2 * n2 + 1], a0, a1, a2, a3, vb[2], vb[3]);
      }
    }
    __syncthreads();   // everyone is done with this buffer before the next iteration's prefetch overwrites it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  uint16_t* ob = out + (size_t)b * N * op + head * 64;
  if (q0 + g < N) {
    uint16_t* r = ob + (size_t)(q0 + g) * op + 2 * t;
#pragma unroll
    for (int n = 0; n < 8; ++n) *reinterpret_cast<uint32_t*>(r + 8 * n) = pack2<DT>(o[n][0] * i0, o[n][1] * i0);
  }
  if (q0 + g + 8 < N) {
    uint16_t* r = ob + (size_t)(q0 + g + 8) * op + 2 * t;
#pragma unroll
    for (int n = 0; n < 8; ++n) *reinterpret_cast<uint32_t*>(r + 8 * n) = pack2<DT>(o[n][2] * i1, o[n][3] * i1);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cross-attention against L <= 8 context tokens: kc / vc fp32 [B][L][C] (projected once per loop), q 16-bit [B][N][C].
// Thread = 8 channels of one token; the 8 threads of a head (64 channels) are 8 consecutive lanes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kXMaxL = 8;

template <int DT>
__global__ void __launch_bounds__(256) cross_attention_small_kernel(const uint16_t* __restrict__ q, int qp,
                                                                    const float* __restrict__ kc,
                                                                    const float* __restrict__ vc, int L, int N, int C,
                                                                    uint16_t* __restrict__ out, int op, float scale,
                                                                    long long total) {
  pdl_prologue();
  const int nvec = C >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // total is padded to a multiple of 32
  const bool valid = i < total;
  const long long tok = valid ? i / nvec : 0;
  const int vi = valid ? (int)(i - tok * nvec) : 0;
  const int b = (int)(tok / N);
  float qf[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(q + tok * qp + vi * 8));
    const float2 f0 = unpack2<DT>(u.x), f1 = unpack2<DT>(u.y), f2 = unpack2<DT>(u.z), f3 = unpack2<DT>(u.w);
    qf[0] = f0.x; qf[1] = f0.y; qf[2] = f1.x; qf[3] = f1.y; qf[4] = f2.x; qf[5] = f2.y; qf[6] = f3.x; qf[7] = f3.y;
  }
  float sc[kXMaxL];
  float mx = -INFINITY;
#pragma unroll
  for (int l = 0; l < kXMaxL; ++l) {
    sc[l] = -INFINITY;
    if (l < L) {
      const float4* kr = reinterpret_cast<const float4*>(kc + ((size_t)b * L + l) * C + vi * 8);
      const float4 k0 = __ldg(kr), k1 = __ldg(kr + 1);
      float d = qf[0] * k0.x + qf[1] * k0.y + qf[2] * k0.z + qf[3] * k0.w + qf[4] * k1.x + qf[5] * k1.y + qf[6] * k1.z +
                qf[7] * k1.w;
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      sc[l] = d * scale;
      mx = fmaxf(mx, sc[l]);
    }
  }
  float den = 0.f, acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
  for (int l = 0; l < kXMaxL; ++l) {
    if (l < L) {
      const float p = expf(sc[l] - mx);
      den += p;
      const float4* vr = reinterpret_cast<const float4*>(vc + ((size_t)b * L + l) * C + vi * 8);
      const float4 v0 = __ldg(vr), v1 = __ldg(vr + 1);
      acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v0.z, acc[2]);
      acc[3] = fmaf(p, v0.w, acc[3]); acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]);
      acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
    }
  }
  if (valid) {
    const float inv = 1.0f / den;
    uint4 o;
    o.x = pack2<DT>(acc[0] * inv, acc[1] * inv);
    o.y = pack2<DT>(acc[2] * inv, acc[3] * inv);
    o.z = pack2<DT>(acc[4] * inv, acc[5] * inv);
    o.w = pack2<DT>(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(out + tok * op + vi * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// y[n][r] = act_out(bias[r] + sum_k w[r][k] * act_in(x[n][k])), fp32, n <= a few dozen rows.  One warp per output row r.
// act: 0 none, 1 SiLU, 2 GELU (exact, erf: transformers' "gelu" of the CLIP text MLP), 3 quick GELU x * sigmoid(1.702 x).
// ---------------------------------------------------------------------------------------------------------------------
D3_DEVINL float act_f(float x, int act) {
  switch (act) {
    case 1: return x / (1.0f + expf(-x));
    case 2: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case 3: return x / (1.0f + expf(-1.702f * x));
    default: return x;
  }
}

__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int n_rows, int K,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int rows, int act_in, int act_out, float* __restrict__ out,
                                                           int out_stride) {
  pdl_prologue();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* wr = w + (size_t)r * K;
  const float bv = bias ? bias[r] : 0.f;
  for (int n0 = 0; n0 < n_rows; n0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kk = lane * 4; kk < K; kk += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + kk));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + j < n_rows) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)(n0 + j) * K + kk));
          acc[j] = fmaf(wv.x, act_f(xv.x, act_in), acc[j]);
          acc[j] = fmaf(wv.y, act_f(xv.y, act_in), acc[j]);
          acc[j] = fmaf(wv.z, act_f(xv.z, act_in), acc[j]);
          acc[j] = fmaf(wv.w, act_f(xv.w, act_in), acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
      if (lane == 0 && n0 + j < n_rows) out[(size_t)(n0 + j) * out_stride + r] = act_f(acc[j] + bv, act_out);
    }
  }
}

// diffusers Timesteps: emb[n] = [sin | cos](t_n * exp(-ln(1e4) * i / (half - shift))), swapped when flip_sin_to_cos
__global__ void timestep_sincos_kernel(const float* __restrict__ t, int nt, int dim, int flip, float shift,
                                       float* __restrict__ out) {
  pdl_prologue();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * half) return;
  const int n = i / half, j = i - n * half;
  const float freq = expf(-9.210340371976184f * (float)j / ((float)half - shift));
  const float arg = t[n] * freq;
  const float sv = sinf(arg), cv = cosf(arg);
  out[(size_t)n * dim + j] = flip ? cv : sv;
  out[(size_t)n * dim + half + j] = flip ? sv : cv;
}

// fp32 NCHW [B][C][H][W] -> 16-bit NHWC channels [c_off, c_off + C) of a [B][H][W][pitch] tensor
template <int DT>
__global__ void __launch_bounds__(256) nchw_to_nhwc16_kernel(const float* __restrict__ x, int B, int C, int H, int W,
                                                             uint16_t* __restrict__ out, int pitch, int c_off) {
  pdl_prologue();
  const long long total = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;                 // b * HW + pixel
    const long long b = p / ((long long)H * W);
    const long long hw = p - b * (long long)H * W;
    const float v = x[(b * C + c) * (long long)H * W + hw];
    out[p * pitch + c_off + c] = (uint16_t)(pack2<DT>(v, 0.f) & 0xFFFFu);
  }
}

// 16-bit NHWC channels [c_off, c_off + C) -> fp32 NCHW
template <int DT>
__global__ void __launch_bounds__(256) nhwc16_to_nchw_kernel(const uint16_t* __restrict__ x, int pitch, int c_off, int B,
                                                             int C, int H, int W, float* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % ((long long)H * W);
    const long long bc = i / ((long long)H * W);
    const int c = (int)(bc % C);
    const long long b = bc / C;
    out[i] = unpack2<DT>((uint32_t)x[(b * H * W + hw) * pitch + c_off + c]).x;
  }
}

// out[b, ho, wo, :] = in[b, floor(ho * H / Ho), floor(wo * W / Wo), :]   (F.interpolate(size=..., mode="nearest"));
// pad_h / pad_w extra zero rows / columns are appended (even-sized copy for the stride-2 convs of odd maps: ho >= Ho or
// wo >= Wo is written as zero)
__global__ void __launch_bounds__(256) resize_nearest_kernel(const uint16_t* __restrict__ in, int B, int H, int W, int C,
                                                             int pin, uint16_t* __restrict__ out, int Ho, int Wo,
                                                             int Hbuf, int Wbuf, int pout) {
  pdl_prologue();
  const int nvec = C >> 3;
  const long long total = (long long)B * Hbuf * Wbuf * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int vi = (int)(i % nvec);
    long long p = i / nvec;
    const int wo = (int)(p % Wbuf);
    p /= Wbuf;
    const int ho = (int)(p % Hbuf);
    const int b = (int)(p / Hbuf);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (ho < Ho && wo < Wo) {
      // torch nearest: src = floor(dst * scale), scale = in / out computed in fp32
      int hs = (int)floorf((float)ho * ((float)H / (float)Ho));
      int ws = (int)floorf((float)wo * ((float)W / (float)Wo));
      if (hs > H - 1) hs = H - 1;
      if (ws > W - 1) ws = W - 1;
      val = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + hs) * W + ws) * pin + vi * 8));
    }
    *reinterpret_cast<uint4*>(out + (((size_t)b * Hbuf + ho) * Wbuf + wo) * pout + vi * 8) = val;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// VAE mid-block attention (one head of 512 channels) runs as two tcgen05 GEMMs per sample (S = Q K^T, O = P V on the
// igemm kernel) with these two helpers in between.
// softmax_rows: in place, P[r][c] = softmax_c(scale * S[r][c]) over the first n_cols columns of a [rows][pitch] 16-bit
// matrix; the padding columns [n_cols, pitch) are written as zero (they are K columns of the second GEMM).
// ---------------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) softmax_rows_kernel(uint16_t* __restrict__ s, int rows, int n_cols, int pitch,
                                                           float scale_log2) {
  pdl_prologue();
  const int row = blockIdx.x;
  if (row >= rows) return;
  uint16_t* r = s + (size_t)row * pitch;
  const int nvec = pitch >> 3;
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(r + vi * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      if (vi * 8 + 2 * k < n_cols) mx = fmaxf(mx, f.x);
      if (vi * 8 + 2 * k + 1 < n_cols) mx = fmaxf(mx, f.y);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
    s_bcast = m;
  }
  __syncthreads();
  mx = s_bcast;
  const float nm = -mx * scale_log2;
  float sum = 0.f;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(r + vi * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      if (vi * 8 + 2 * k < n_cols) sum += exp2f(fmaf(f.x, scale_log2, nm));
      if (vi * 8 + 2 * k + 1 < n_cols) sum += exp2f(fmaf(f.y, scale_log2, nm));
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  __syncthreads();   // s_red is reused
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    s_bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = s_bcast;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(r + vi * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack2<DT>(w[k]);
      const float p0 = (vi * 8 + 2 * k < n_cols) ? exp2f(fmaf(f.x, scale_log2, nm)) * inv : 0.f;
      const float p1 = (vi * 8 + 2 * k + 1 < n_cols) ? exp2f(fmaf(f.y, scale_log2, nm)) * inv : 0.f;
      o[k] = pack2<DT>(p0, p1);
    }
    *reinterpret_cast<uint4*>(r + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[c][n] = in[n][c] for n < N, 0 for N <= n < Np   (V -> V^T as the K-major weight operand of O = P V)
__global__ void __launch_bounds__(256) transpose_pad_kernel(const uint16_t* __restrict__ in, int N, int C, int pin,
                                                            uint16_t* __restrict__ out, int Np) {
  pdl_prologue();
  __shared__ uint16_t tile[32][34];
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + tx;
    tile[i][tx] = (n < N && c < C) ? in[(size_t)n * pin + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    if (c < C && n < Np) out[(size_t)c * Np + n] = tile[tx][i];
  }
}

// 1x1 conv over fp32 NCHW with <= 8 channels on both sides (quant_conv 8 -> 8, post_quant_conv 4 -> 4 of AutoencoderKL)
__global__ void __launch_bounds__(256) conv1x1_nchw_small_kernel(const float* __restrict__ x, int Cin, int Cout,
                                                                 long long HW, long long total,
                                                                 const float* __restrict__ w, const float* __restrict__ b,
                                                                 float* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < Cin ? x[(n * Cin + c) * HW + hw] : 0.f;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < Cout) {
        float acc = b ? b[o] : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < Cin) acc = fmaf(w[o * Cin + c], v[c], acc);
        out[(n * Cout + o) * HW + hw] = acc;
      }
    }
  }
}

int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace d3

using namespace d3;

#define D3_16BIT(dtype) D3_REQUIRE((dtype) == D3_DTYPE_BF16 || (dtype) == D3_DTYPE_F16, "16-bit dtypes only (got %d)", (dtype))

extern "C" int d3_gn_stats(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2, int B, int HW,
                           double* chan_stats, int dtype, void* stream) {
  D3_16BIT(dtype);
  GnGenArgs a{};
  dim3 grid;
  int threads = 0;
  if (int rc = gn_gen_setup(a, x1, C1, pitch1, x2, C2, pitch2, B, HW, chan_stats, &grid, &threads)) return rc;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_stats_kernel<kF16>, grid, dim3(threads), 0, st, a));
  else D3_CHECK_CUDA(launch_pdl(gn_stats_kernel<kBF16>, grid, dim3(threads), 0, st, a));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_gn_apply(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2,
                           const double* chan_stats, const float* gamma, const float* beta, int B, int HW, int groups,
                           float eps, int apply_silu, void* out, int out_pitch, int dtype, void* stream) {
  D3_16BIT(dtype);
  const int C = C1 + C2;
  D3_REQUIRE(gamma && beta && out && out_pitch % 8 == 0 && groups >= 1 && groups <= 64 && C % groups == 0,
             "gn_apply: bad arguments (C=%d groups=%d)", C, groups);
  GnGenArgs a{};
  dim3 grid;
  int threads = 0;
  if (int rc = gn_gen_setup(a, x1, C1, pitch1, x2, C2, pitch2, B, HW, const_cast<double*>(chan_stats), &grid, &threads))
    return rc;
  a.gamma = gamma;
  a.beta = beta;
  a.out = static_cast<uint16_t*>(out);
  a.pout = out_pitch;
  a.groups = groups;
  a.silu = apply_silu;
  a.eps = eps;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_apply_gen_kernel<kF16>, grid, dim3(threads), 0, st, a));
  else D3_CHECK_CUDA(launch_pdl(gn_apply_gen_kernel<kBF16>, grid, dim3(threads), 0, st, a));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_layernorm(const void* x, int x_pitch, const float* gamma, const float* beta, int64_t M, int C,
                            float eps, void* out, int out_pitch, int dtype, void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(x && gamma && beta && out && C % 8 == 0 && C <= kLnMaxV * 256 && x_pitch % 8 == 0 && out_pitch % 8 == 0 &&
                 M > 0,
             "layernorm: unsupported C=%d (multiple of 8, <= %d)", C, kLnMaxV * 256);
  const int blocks = (int)((M * 32 + 255) / 256);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(layernorm_kernel<kF16>, dim3(blocks), dim3(256), 0, st, static_cast<const uint16_t*>(x),
                             x_pitch, gamma, beta, static_cast<uint16_t*>(out), out_pitch, (long long)M, C, eps));
  else
    D3_CHECK_CUDA(launch_pdl(layernorm_kernel<kBF16>, dim3(blocks), dim3(256), 0, st, static_cast<const uint16_t*>(x),
                             x_pitch, gamma, beta, static_cast<uint16_t*>(out), out_pitch, (long long)M, C, eps));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_geglu(const void* x, int x_pitch, int64_t M, int D, void* out, int out_pitch, int dtype,
                        void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(x && out && D % 8 == 0 && x_pitch % 8 == 0 && out_pitch % 8 == 0 && M > 0, "geglu: bad arguments");
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = grid_for(M * (D / 8), 256);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(geglu_kernel<kF16>, dim3(blocks), dim3(256), 0, st, static_cast<const uint16_t*>(x), x_pitch,
                             static_cast<uint16_t*>(out), out_pitch, (long long)M, D));
  else
    D3_CHECK_CUDA(launch_pdl(geglu_kernel<kBF16>, dim3(blocks), dim3(256), 0, st, static_cast<const uint16_t*>(x), x_pitch,
                             static_cast<uint16_t*>(out), out_pitch, (long long)M, D));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_attention_hd64(const void* q, int q_pitch, const void* k, int k_pitch, const void* v, int v_pitch,
                                 int B, int N, int heads, void* out, int out_pitch, int dtype, void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(q && k && v && out && B > 0 && N > 0 && heads > 0 && q_pitch % 8 == 0 && k_pitch % 8 == 0 &&
                 v_pitch % 8 == 0 && out_pitch % 2 == 0,
             "attention_hd64: bad arguments");
  D3_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) % 16 == 0, "attention_hd64: pointers must be 16-byte aligned");
  const float scale_log2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  dim3 grid(ceil_div(N, 64), heads, B);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(attention_hd64_kernel<kF16>, grid, dim3(128), 0, st, static_cast<const uint16_t*>(q), q_pitch,
                             static_cast<const uint16_t*>(k), k_pitch, static_cast<const uint16_t*>(v), v_pitch, N,
                             static_cast<uint16_t*>(out), out_pitch, scale_log2));
  else
    D3_CHECK_CUDA(launch_pdl(attention_hd64_kernel<kBF16>, grid, dim3(128), 0, st, static_cast<const uint16_t*>(q),
                             q_pitch, static_cast<const uint16_t*>(k), k_pitch, static_cast<const uint16_t*>(v), v_pitch,
                             N, static_cast<uint16_t*>(out), out_pitch, scale_log2));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_cross_attention_small(const void* q, int q_pitch, const float* kc, const float* vc, int B, int N,
                                        int L, int C, void* out, int out_pitch, int dtype, void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(q && kc && vc && out && L >= 1 && L <= kXMaxL && C % 64 == 0 && q_pitch % 8 == 0 && out_pitch % 8 == 0,
             "cross_attention_small: bad arguments (L=%d <= %d, C=%d multiple of 64)", L, kXMaxL, C);
  const long long total = (long long)B * N * (C / 8);
  const int blocks = (int)((total + 255) / 256);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(cross_attention_small_kernel<kF16>, dim3(blocks), dim3(256), 0, st,
                             static_cast<const uint16_t*>(q), q_pitch, kc, vc, L, N, C, static_cast<uint16_t*>(out),
                             out_pitch, 0.125f, total));
  else
    D3_CHECK_CUDA(launch_pdl(cross_attention_small_kernel<kBF16>, dim3(blocks), dim3(256), 0, st,
                             static_cast<const uint16_t*>(q), q_pitch, kc, vc, L, N, C, static_cast<uint16_t*>(out),
                             out_pitch, 0.125f, total));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_linear_small(const float* x, int n_rows, int K, const float* w, const float* bias, int rows,
                               int act_in, int act_out, float* out, int out_stride, void* stream) {
  D3_REQUIRE(x && w && out && n_rows > 0 && rows > 0 && K % 4 == 0 && out_stride >= rows, "linear_small: bad arguments");
  D3_REQUIRE(act_in >= 0 && act_in <= 3 && act_out >= 0 && act_out <= 3, "linear_small: activation code must be 0..3");
  const int blocks = ceil_div(rows * 32, 256);
  D3_CHECK_CUDA(launch_pdl(linear_small_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, n_rows,
                           K, w, bias, rows, act_in, act_out, out, out_stride));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_timestep_sincos(const float* t, int nt, int dim, int flip_sin_to_cos, float freq_shift, float* out,
                                  void* stream) {
  D3_REQUIRE(t && out && nt > 0 && dim > 0 && dim % 2 == 0, "timestep_sincos: bad arguments");
  const int n = nt * (dim / 2);
  D3_CHECK_CUDA(launch_pdl(timestep_sincos_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream),
                           t, nt, dim, flip_sin_to_cos, freq_shift, out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_nchw_to_nhwc16(const float* x, int B, int C, int H, int W, void* out, int out_pitch, int c_off,
                                 int dtype, void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(x && out && c_off >= 0 && c_off + C <= out_pitch, "nchw_to_nhwc16: bad arguments");
  const long long total = (long long)B * C * H * W;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(nchw_to_nhwc16_kernel<kF16>, dim3(grid_for(total, 256)), dim3(256), 0, st, x, B, C, H, W,
                             static_cast<uint16_t*>(out), out_pitch, c_off));
  else
    D3_CHECK_CUDA(launch_pdl(nchw_to_nhwc16_kernel<kBF16>, dim3(grid_for(total, 256)), dim3(256), 0, st, x, B, C, H, W,
                             static_cast<uint16_t*>(out), out_pitch, c_off));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_nhwc16_to_nchw(const void* x, int pitch, int c_off, int B, int C, int H, int W, float* out, int dtype,
                                 void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(x && out && c_off >= 0 && c_off + C <= pitch, "nhwc16_to_nchw: bad arguments");
  const long long total = (long long)B * C * H * W;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(nhwc16_to_nchw_kernel<kF16>, dim3(grid_for(total, 256)), dim3(256), 0, st,
                             static_cast<const uint16_t*>(x), pitch, c_off, B, C, H, W, out));
  else
    D3_CHECK_CUDA(launch_pdl(nhwc16_to_nchw_kernel<kBF16>, dim3(grid_for(total, 256)), dim3(256), 0, st,
                             static_cast<const uint16_t*>(x), pitch, c_off, B, C, H, W, out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_resize_nearest_nhwc16(const void* in, int B, int H, int W, int C, int in_pitch, void* out, int Ho,
                                        int Wo, int Hbuf, int Wbuf, int out_pitch, void* stream) {
  D3_REQUIRE(in && out && C % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0 && Hbuf >= Ho && Wbuf >= Wo,
             "resize_nearest: bad arguments");
  const long long total = (long long)B * Hbuf * Wbuf * (C / 8);
  D3_CHECK_CUDA(launch_pdl(resize_nearest_kernel, dim3(grid_for(total, 256)), dim3(256), 0,
                           static_cast<cudaStream_t>(stream), static_cast<const uint16_t*>(in), B, H, W, C, in_pitch,
                           static_cast<uint16_t*>(out), Ho, Wo, Hbuf, Wbuf, out_pitch));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_softmax_rows(void* s, int rows, int n_cols, int pitch, float scale, int dtype, void* stream) {
  D3_16BIT(dtype);
  D3_REQUIRE(s && rows > 0 && n_cols > 0 && n_cols <= pitch && pitch % 8 == 0, "softmax_rows: bad arguments");
  const float sl2 = scale * 1.4426950408889634f;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(softmax_rows_kernel<kF16>, dim3(rows), dim3(256), 0, st, static_cast<uint16_t*>(s), rows,
                             n_cols, pitch, sl2));
  else
    D3_CHECK_CUDA(launch_pdl(softmax_rows_kernel<kBF16>, dim3(rows), dim3(256), 0, st, static_cast<uint16_t*>(s), rows,
                             n_cols, pitch, sl2));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_transpose_pad16(const void* in, int N, int C, int in_pitch, void* out, int Np, void* stream) {
  D3_REQUIRE(in && out && N > 0 && C > 0 && Np >= N, "transpose_pad16: bad arguments");
  dim3 grid(ceil_div(Np, 32), ceil_div(C, 32));
  D3_CHECK_CUDA(launch_pdl(transpose_pad_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
                           static_cast<const uint16_t*>(in), N, C, in_pitch, static_cast<uint16_t*>(out), Np));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_conv1x1_nchw_small(const float* x, int B, int Cin, int Cout, int64_t HW, const float* w,
                                     const float* bias, float* out, void* stream) {
  D3_REQUIRE(x && w && out && Cin >= 1 && Cin <= 8 && Cout >= 1 && Cout <= 8 && B > 0 && HW > 0,
             "conv1x1_nchw_small: channel counts must be in [1, 8]");
  const long long total = (long long)B * HW;
  D3_CHECK_CUDA(launch_pdl(conv1x1_nchw_small_kernel, dim3(grid_for(total, 256)), dim3(256), 0,
                           static_cast<cudaStream_t>(stream), x, Cin, Cout, (long long)HW, total, w, bias, out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_groupnorm_one_pass(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2,
                                     const float* gamma, const float* beta, int B, int HW, int groups, float eps,
                                     int apply_silu, void* out, int out_pitch, int dtype, void* stream) {
  D3_16BIT(dtype);
  const int C = C1 + C2;
  D3_REQUIRE(x1 && gamma && beta && out && C1 > 0 && C1 % 8 == 0 && C2 % 8 == 0 && pitch1 % 8 == 0 && out_pitch % 8 == 0 &&
                 (C2 == 0 || (x2 && pitch2 % 8 == 0)) && groups >= 1 && C % groups == 0,
             "groupnorm_one_pass: bad arguments (C=%d groups=%d)", C, groups);
  const int cpg = C / groups;
  D3_REQUIRE(cpg % 2 == 0, "groupnorm_one_pass: channels per group must be even (got %d)", cpg);
  int gpb = 0;
  for (int g = 1; g <= 8; ++g)
    if ((g * cpg) % 8 == 0 && groups % g == 0) { gpb = g; break; }
  D3_REQUIRE(gpb > 0, "groupnorm_one_pass: no block shape for %d channels per group", cpg);
  GnOneArgs a{static_cast<const uint16_t*>(x1), static_cast<const uint16_t*>(x2), C1, C2, pitch1, pitch2, HW, cpg, gpb,
              gamma, beta, static_cast<uint16_t*>(out), out_pitch, apply_silu, eps};
  dim3 grid(groups / gpb, B);
  const long long work = (long long)HW * gpb * cpg / 8;
  const int threads = work >= 4096 ? 512 : (work >= 1024 ? 256 : 128);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == D3_DTYPE_F16) D3_CHECK_CUDA(launch_pdl(gn_one_kernel<kF16>, grid, dim3(threads), 0, st, a));
  else D3_CHECK_CUDA(launch_pdl(gn_one_kernel<kBF16>, grid, dim3(threads), 0, st, a));
  D3_LAUNCH_CHECK();
  return D3_OK;
}
