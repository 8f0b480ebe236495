// fp32 "reference-precision" variants of the UNet kernels (D3_DTYPE_F32): same entry points, same NHWC layout, same
// quad statistics, but fp32 storage and CUDA-core FFMA arithmetic with fp32 accumulation.
//
// Why they exist: BASELINE config 4 (SceneFlow-shape 544x960, DDPM 100 steps) is a *tolerance* configuration: per-step
// model-output max-abs <= 1e-4 against the fp32 reference.  tcgen05 has no fp32 MMA (kind::tf32 keeps 10 mantissa
// bits, fewer than fp16), so that bound is met on the CUDA cores; throughput is not the point of this path
// (~1/40 of the tensor-core path).  Everything else (pipeline, schedulers, graphs, sharding) is shared.
//
// conv: implicit GEMM, 128 pixels x 128 output channels per CTA, K-blocks of 16 channels of one filter tap,
// 256 threads x (8 x 8) register tile, same operand semantics as d3_conv2d_igemm (taps incl. sub-pixel upsample
// phases, stride 2, folded 1x1 shortcut slabs, identity residual, per-(sample, quad) statistics).
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kF32Threads = 256;

struct ConvF32Args {
  const float* x;
  const float* sc1;
  const float* sc2;
  const float* w;  // [Cout][Ktot]
  const float* bias;
  const float* residual;
  float* out;
  double* stats;
  int B, H, W, Ho, Wo, Cin, C1, C2, Cout, Ktot;
  int x_pitch, sc1_pitch, sc2_pitch, res_pitch, out_pitch, bias_bstride;
  int stride, ntaps;
  signed char tap_dh[9], tap_dw[9];
  int TH, TW, tilesH, tilesW;
  // output addressing (sub-pixel phase writes a strided view of the 2x larger tensor)
  long long out_off;      // element offset of pixel (0,0) of this launch's view
  long long out_sw, out_sh, out_sb;  // element strides between output pixels in w, h, b
};

__global__ void __launch_bounds__(kF32Threads) conv2d_f32_kernel(const ConvF32Args a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_stat[BN / 4][2];
  const int tid = threadIdx.x;
  const int nt = blockIdx.y;
  int p = blockIdx.x;
  const int tw_i = p % a.tilesW;
  p /= a.tilesW;
  const int th_i = p % a.tilesH;
  const int b = p / a.tilesH;
  const int h0 = th_i * a.TH, w0 = tw_i * a.TW;
  const int n0 = nt * BN;
  // loader mapping: thread -> (row m = tid / 2, half = tid % 2 -> 8 consecutive k)
  const int lm = tid >> 1, lk = (tid & 1) * 8;
  const int lth = lm / a.TW, ltw = lm - lth * a.TW;
  const int oh = h0 + lth, ow = w0 + ltw;
  const bool lvalid = (lth < a.TH) && (oh < a.Ho) && (ow < a.Wo);
  // compute mapping: 16 x 16 threads, each 8 rows (pixels) x 8 cols (channels)
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int kbMain = a.ntaps * (a.Cin / BK);
  const int kbS1 = a.C1 / BK, kbS2 = a.C2 / BK;
  const int kbTot = kbMain + kbS1 + kbS2;
  for (int kb = 0; kb < kbTot; ++kb) {
    // ---- gather A (pixels x 16 k) and B (16 k x 128 couts) into shared memory
    float av[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) av[e] = 0.f;
    if (lvalid) {
      const float* src = nullptr;
      if (kb < kbMain) {
        const int tap = kb / (a.Cin / BK);
        const int c0 = (kb - tap * (a.Cin / BK)) * BK;
        const int ih = oh * a.stride + a.tap_dh[tap], iw = ow * a.stride + a.tap_dw[tap];
        if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
          src = a.x + (((size_t)b * a.H + ih) * a.W + iw) * a.x_pitch + c0 + lk;
      } else if (kb < kbMain + kbS1) {
        src = a.sc1 + (((size_t)b * a.Ho + oh) * a.Wo + ow) * a.sc1_pitch + (kb - kbMain) * BK + lk;
      } else {
        src = a.sc2 + (((size_t)b * a.Ho + oh) * a.Wo + ow) * a.sc2_pitch + (kb - kbMain - kbS1) * BK + lk;
      }
      if (src) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
        av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
      }
    }
    float bv[8];
    {
      const float* wsrc = a.w + (size_t)(n0 + lm) * a.Ktot + (size_t)kb * BK + lk;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(wsrc));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(wsrc) + 1);
      bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
      bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
    }
    __syncthreads();  // previous iteration's reads are done
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      As[lk + e][lm] = av[e];
      Bs[lk + e][lm] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float ar[8], br[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      ar[0] = a0.x; ar[1] = a0.y; ar[2] = a0.z; ar[3] = a0.w; ar[4] = a1.x; ar[5] = a1.y; ar[6] = a1.z; ar[7] = a1.w;
      br[0] = b0.x; br[1] = b0.y; br[2] = b0.z; br[3] = b0.w; br[4] = b1.x; br[5] = b1.y; br[6] = b1.z; br[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  // ---- epilogue
  if (tid < BN / 4) s_stat[tid][0] = s_stat[tid][1] = 0.f;
  __syncthreads();
  const int cbase = n0 + tx * 8;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = a.bias[(size_t)b * a.bias_bstride + cbase + j];
  float q_s[2] = {0.f, 0.f}, q_q[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = ty * 8 + i;
    const int mth = m / a.TW, mtw = m - mth * a.TW;
    const int yy = h0 + mth, xx = w0 + mtw;
    if (mth < a.TH && yy < a.Ho && xx < a.Wo) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bsum[j];
      if (a.residual) {
        const float* rp = a.residual + (((size_t)b * a.Ho + yy) * a.Wo + xx) * a.res_pitch + cbase;
        const float4 r0 = *reinterpret_cast<const float4*>(rp), r1 = *reinterpret_cast<const float4*>(rp + 4);
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
      float* op = a.out + a.out_off + (long long)b * a.out_sb + (long long)yy * a.out_sh + (long long)xx * a.out_sw + cbase;
      *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        q_s[j >> 2] += v[j];
        q_q[j >> 2] = fmaf(v[j], v[j], q_q[j >> 2]);
      }
    }
  }
  if (a.stats) {
    atomicAdd(&s_stat[tx * 2][0], q_s[0]);
    atomicAdd(&s_stat[tx * 2][1], q_q[0]);
    atomicAdd(&s_stat[tx * 2 + 1][0], q_s[1]);
    atomicAdd(&s_stat[tx * 2 + 1][1], q_q[1]);
    __syncthreads();
    if (tid < BN / 4) {
      double* dst = a.stats + ((size_t)b * (a.Cout >> 2) + (n0 >> 2) + tid) * 2;
      atomicAdd(dst, (double)s_stat[tid][0]);
      atomicAdd(dst + 1, (double)s_stat[tid][1]);
    }
  }
}

// ---- GroupNorm (+SiLU) apply, fp32, up to two concatenated sources (same contract as gn_apply_kernel) -----------
struct GnF32Args {
  const float* x1;
  const float* x2;
  const double* st1;
  const double* st2;
  const float* gamma;
  const float* beta;
  float* out;
  int C1, C2, p1, p2, pout, HW, groups, silu, pix_per_block;
  float eps;
};

__global__ void __launch_bounds__(256) gn_apply_f32_kernel(const GnF32Args a) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.y;
  const int C = a.C1 + a.C2;
  const int cpg = C / a.groups, qpg = cpg >> 2;
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
  __syncthreads();
  const int nvec = C >> 2, nvec1 = a.C1 >> 2;  // float4 vectors per pixel
  const int vec = threadIdx.x % nvec, prow = threadIdx.x / nvec, pstride = blockDim.x / nvec;
  float sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = vec * 4 + k, g = c / cpg;
    sc[k] = a.gamma[c] * s_rstd[g];
    sh[k] = a.beta[c] - s_mean[g] * sc[k];
  }
  const int pix0 = blockIdx.x * a.pix_per_block;
  int npix = a.HW - pix0;
  if (npix > a.pix_per_block) npix = a.pix_per_block;
  const size_t base = (size_t)b * a.HW + pix0;
  const float* src = (vec < nvec1) ? a.x1 + base * a.p1 + vec * 4 : a.x2 + base * a.p2 + (vec - nvec1) * 4;
  const int sp = (vec < nvec1) ? a.p1 : a.p2;
  float* dst = a.out + base * a.pout + vec * 4;
  for (int p = prow; p < npix; p += pstride) {
    const float4 v = *reinterpret_cast<const float4*>(src + (size_t)p * sp);
    float y[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
    if (a.silu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.0f + expf(-y[k]));
    }
    *reinterpret_cast<float4*>(dst + (size_t)p * a.pout) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

// ---- attention, head_dim 8, fp32 I/O: one query per thread, K/V of one (sample, head) in shared memory ----------
__global__ void __launch_bounds__(256) attention_hd8_f32_kernel(const float* __restrict__ qkv, int pitch, int C, int N,
                                                                float* __restrict__ out, int opitch) {
  extern __shared__ float4 s_kv[];
  float4* s_k = s_kv;
  float4* s_v = s_kv + 2 * N;
  const int head = blockIdx.y, b = blockIdx.z;
  const float* base = qkv + (size_t)b * N * pitch + head * 8;
  for (int j = threadIdx.x; j < N; j += 256) {
    const float4* kp = reinterpret_cast<const float4*>(base + (size_t)j * pitch + C);
    const float4* vp = reinterpret_cast<const float4*>(base + (size_t)j * pitch + 2 * C);
    s_k[2 * j] = kp[0]; s_k[2 * j + 1] = kp[1];
    s_v[2 * j] = vp[0]; s_v[2 * j + 1] = vp[1];
  }
  __syncthreads();
  const int qi = blockIdx.x * 256 + threadIdx.x;
  if (qi >= N) return;
  const float sc = 0.35355339059327373f;
  const float4 q0 = *reinterpret_cast<const float4*>(base + (size_t)qi * pitch);
  const float4 q1 = *reinterpret_cast<const float4*>(base + (size_t)qi * pitch + 4);
  float m = -INFINITY, l = 0.f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < N; ++j) {
    const float4 k0 = s_k[2 * j], k1 = s_k[2 * j + 1];
    float s = q0.x * k0.x;
    s = fmaf(q0.y, k0.y, s); s = fmaf(q0.z, k0.z, s); s = fmaf(q0.w, k0.w, s);
    s = fmaf(q1.x, k1.x, s); s = fmaf(q1.y, k1.y, s); s = fmaf(q1.z, k1.z, s); s = fmaf(q1.w, k1.w, s);
    s *= sc;
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), pj = expf(s - mn);
    m = mn;
    l = l * corr + pj;
    const float4 v0 = s_v[2 * j], v1 = s_v[2 * j + 1];
    o[0] = fmaf(pj, v0.x, o[0] * corr); o[1] = fmaf(pj, v0.y, o[1] * corr);
    o[2] = fmaf(pj, v0.z, o[2] * corr); o[3] = fmaf(pj, v0.w, o[3] * corr);
    o[4] = fmaf(pj, v1.x, o[4] * corr); o[5] = fmaf(pj, v1.y, o[5] * corr);
    o[6] = fmaf(pj, v1.z, o[6] * corr); o[7] = fmaf(pj, v1.w, o[7] * corr);
  }
  const float inv = 1.0f / l;
  float* op = out + ((size_t)b * N + qi) * opitch + head * 8;
  *reinterpret_cast<float4*>(op) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
  *reinterpret_cast<float4*>(op + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
}

// ---- conv_norm_out + SiLU + conv_out (128 -> dc), fp32: one output pixel per thread ---------------------------------
__global__ void __launch_bounds__(128) conv_out_f32_kernel(const float* __restrict__ x, int pitch,
                                                           const double* __restrict__ stats,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int B, int H, int W, int dc,
                                                           float eps, float* __restrict__ out) {
  __shared__ float s_sc[128], s_sh[128];
  const int b = blockIdx.z;
  if (threadIdx.x < 128) {
    const int c = threadIdx.x, g = c >> 2;
    const double s = stats[((size_t)b * 32 + g) * 2], ss = stats[((size_t)b * 32 + g) * 2 + 1];
    const double n = (double)H * W * 4.0;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    s_sc[c] = gamma[c] * rstd;
    s_sh[c] = beta[c] - (float)mean * s_sc[c];
  }
  __syncthreads();
  const int xx = blockIdx.x * 128 + threadIdx.x, yy = blockIdx.y;
  if (xx >= W) return;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int tap = 0; tap < 9; ++tap) {
    const int ih = yy + tap / 3 - 1, iw = xx + tap % 3 - 1;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
    const float* px = x + (((size_t)b * H + ih) * W + iw) * pitch;
    for (int c = 0; c < 128; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(px + c);
      float y[4] = {fmaf(v.x, s_sc[c], s_sh[c]), fmaf(v.y, s_sc[c + 1], s_sh[c + 1]),
                    fmaf(v.z, s_sc[c + 2], s_sh[c + 2]), fmaf(v.w, s_sc[c + 3], s_sh[c + 3])};
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.0f + expf(-y[k]));
      for (int d = 0; d < dc; ++d) {
        const float* wp = w + ((size_t)d * 9 + tap) * 128 + c;
        acc[d] = fmaf(y[0], wp[0], acc[d]); acc[d] = fmaf(y[1], wp[1], acc[d]);
        acc[d] = fmaf(y[2], wp[2], acc[d]); acc[d] = fmaf(y[3], wp[3], acc[d]);
      }
    }
  }
  for (int d = 0; d < dc; ++d) out[(((size_t)b * dc + d) * H + yy) * W + xx] = acc[d] + bias[d];
}

// ---- conv_in front end, fp32: same [B,H,W,128] im2col row layout as im2col_input_kernel ------------------------------
__global__ void __launch_bounds__(256) im2col_input_f32_kernel(const float* __restrict__ xs, int Cs,
                                                               const float* __restrict__ cond, int Cc, int B, int H,
                                                               int W, float* __restrict__ out) {
  const int Cin = Cs + Cc, K = 9 * Cin;
  const size_t total = (size_t)B * H * W * 128;
  const size_t HWs = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i & 127);
    const size_t p = i >> 7;
    const int w = (int)(p % W);
    const size_t t = p / W;
    const int h = (int)(t % H);
    const size_t b = t / H;
    float v = 0.f;
    if (k < K) {
      const int tap = k / Cin, c = k - tap * Cin;
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = (c < Cs) ? xs[(b * Cs + c) * HWs + (size_t)hh * W + ww] : cond[(b * Cc + (c - Cs)) * HWs + (size_t)hh * W + ww];
    }
    out[i] = v;
  }
}

void choose_tile_f32(int Ho, int Wo, int* TH, int* TW) {
  double best = -1.0;
  int bh = 1, bw = 1;
  for (int tw = 1; tw <= 128 && tw <= Wo; ++tw) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th < 1) continue;
    const double tiles = double(ceil_div(Ho, th)) * ceil_div(Wo, tw);
    const double eff = double(Ho) * Wo / (tiles * 128.0);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > bw)) {
      best = eff;
      bh = th;
      bw = tw;
    }
  }
  *TH = bh;
  *TW = bw;
}

}  // namespace

int conv2d_f32(const d3_conv_params* p, cudaStream_t stream) {
  D3_REQUIRE(p->Cin % 16 == 0 && p->C1 % 16 == 0 && p->C2 % 16 == 0 && p->Cout % 128 == 0,
             "conv(f32): channel counts must be multiples of 16 (Cout of 128)");
  D3_REQUIRE(p->x_pitch % 4 == 0 && p->out_pitch % 4 == 0, "conv(f32): pitches must be multiples of 4");
  const int phase = p->up2x_phase;
  ConvF32Args a{};
  a.x = static_cast<const float*>(p->x);
  a.sc1 = static_cast<const float*>(p->sc1);
  a.sc2 = static_cast<const float*>(p->sc2);
  a.w = static_cast<const float*>(p->w);
  a.bias = p->bias;
  a.residual = static_cast<const float*>(p->residual);
  a.out = static_cast<float*>(p->out);
  a.stats = p->stats;
  a.B = p->B; a.H = p->H; a.W = p->W;
  a.Ho = p->H / p->stride; a.Wo = p->W / p->stride;
  a.Cin = p->Cin; a.C1 = p->sc1 ? p->C1 : 0; a.C2 = p->sc2 ? p->C2 : 0; a.Cout = p->Cout;
  a.x_pitch = p->x_pitch; a.sc1_pitch = p->sc1_pitch; a.sc2_pitch = p->sc2_pitch; a.res_pitch = p->res_pitch;
  a.out_pitch = p->out_pitch; a.bias_bstride = p->bias_bstride;
  a.stride = p->stride;
  const int pad = p->ksize / 2;
  if (phase) {
    const int py = (phase - 1) / 2, px = (phase - 1) % 2;
    a.ntaps = 4;
    for (int t = 0; t < 4; ++t) {
      a.tap_dh[t] = (signed char)(py - 1 + t / 2);
      a.tap_dw[t] = (signed char)(px - 1 + t % 2);
    }
    a.out_off = ((long long)py * (2 * a.Wo) + px) * p->out_pitch;
    a.out_sw = 2LL * p->out_pitch;
    a.out_sh = 2LL * (2 * a.Wo) * p->out_pitch;
    a.out_sb = (long long)(2 * a.Ho) * (2 * a.Wo) * p->out_pitch;
  } else {
    a.ntaps = p->ksize * p->ksize;
    for (int t = 0; t < a.ntaps; ++t) {
      a.tap_dh[t] = (signed char)(t / p->ksize - pad);
      a.tap_dw[t] = (signed char)(t % p->ksize - pad);
    }
    a.out_off = 0;
    a.out_sw = p->out_pitch;
    a.out_sh = (long long)a.Wo * p->out_pitch;
    a.out_sb = (long long)a.Ho * a.Wo * p->out_pitch;
  }
  a.Ktot = a.ntaps * a.Cin + a.C1 + a.C2;
  choose_tile_f32(a.Ho, a.Wo, &a.TH, &a.TW);
  a.tilesH = ceil_div(a.Ho, a.TH);
  a.tilesW = ceil_div(a.Wo, a.TW);
  dim3 grid(p->B * a.tilesH * a.tilesW, p->Cout / BN);
  conv2d_f32_kernel<<<grid, kF32Threads, 0, stream>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

int groupnorm_f32(const void* x1, int C1, int p1, const double* stats1, const void* x2, int C2, int p2,
                  const double* stats2, const float* gamma, const float* beta, int B, int HW, int groups, float eps,
                  int apply_silu, void* out, int pout, cudaStream_t stream) {
  const int C = C1 + C2;
  D3_REQUIRE(C <= 1024 && C % groups == 0 && (C / groups) % 4 == 0 && groups <= 32 && C1 % 4 == 0 && C2 % 4 == 0,
             "groupnorm(f32): unsupported C=%d", C);
  GnF32Args a{static_cast<const float*>(x1), static_cast<const float*>(x2), stats1, stats2, gamma, beta,
              static_cast<float*>(out), C1, C2, p1, p2, pout, HW, groups, apply_silu, 0, eps};
  int blocks_per_sample = ceil_div(num_sms() * 8, B);
  int ppb = ceil_div(HW, blocks_per_sample);
  if (ppb < 16) ppb = 16;
  a.pix_per_block = ppb;
  const int nvec = C / 4;
  const int threads = nvec * (256 / nvec > 0 ? 256 / nvec : 1);
  dim3 grid(ceil_div(HW, ppb), B);
  gn_apply_f32_kernel<<<grid, threads, 0, stream>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

int attention_f32(const void* qkv, int pitch, int B, int N, int C, void* out, int opitch, cudaStream_t stream) {
  const size_t smem = (size_t)N * 64;
  D3_REQUIRE(smem <= 220 * 1024, "attention(f32): N=%d exceeds the shared-memory K/V budget", N);
  D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attention_hd8_f32_kernel), smem));
  dim3 grid(ceil_div(N, 256), C / 8, B);
  attention_hd8_f32_kernel<<<grid, 256, smem, stream>>>(static_cast<const float*>(qkv), pitch, C, N,
                                                        static_cast<float*>(out), opitch);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

int im2col_f32(const float* sample, int Cs, const float* cond, int Cc, int B, int H, int W, void* out,
               cudaStream_t stream) {
  im2col_input_f32_kernel<<<num_sms() * 16, 256, 0, stream>>>(sample, Cs, cond, Cc, B, H, W, static_cast<float*>(out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

int conv_out_f32(const void* x, int pitch, const double* stats, const float* gamma, const float* beta, const float* w,
                 const float* bias, int B, int H, int W, int dc, float eps, float* out, cudaStream_t stream) {
  dim3 grid(ceil_div(W, 128), H, B);
  conv_out_f32_kernel<<<grid, 128, 0, stream>>>(static_cast<const float*>(x), pitch, stats, gamma, beta, w, bias, B, H,
                                                W, dc, eps, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

}  // namespace d3
