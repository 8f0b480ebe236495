// The two "edge" convolutions of the UNet, whose GEMM shapes are degenerate:
//
//  conv_in  (Cin = 5/8/12 -> 128, K = 9*Cin <= 126): the fp32 NCHW model input (noisy sample + conditioning,
//           custom_pipelines.py:158-171) is gathered straight into a bf16 im2col matrix [B,H,W,128]
//           (k = tap*Cin + c, zero-padded to 128) by d3_im2col_input; the tcgen05 igemm then runs it as a 1x1
//           conv with K = 128.  This fuses torch.cat + dtype cast + NCHW->NHWC + zero padding into one pass.
//
//  conv_out (128 -> dc, N = 1 or 3): GroupNorm + SiLU + 3x3 conv fused in one CUDA-core kernel: a (8+2)x(32+2)x128
//           input region is read once, normalised + SiLU'd into shared memory, then every thread produces one output
//           pixel (9 taps x 128 ch) and writes the fp32 NCHW model output the schedulers consume.
//           Algorithmic bytes: (128 + dc*2) * P * 2 B per sample.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "sched_math.cuh"

namespace d3 {
namespace {

// ------------------------------------------------------------------------------------------------
// im2col gather for conv_in
// ------------------------------------------------------------------------------------------------
constexpr int kI2cPix = 64;     // pixels (of one image row) per block
constexpr int kI2cPitchW = 65;  // smem words per pixel row (64 + 1 pad: conflict-free column writes AND row reads)

template <int DT>
__global__ void __launch_bounds__(256) im2col_input_kernel(const float* __restrict__ xs, int Cs,
                                                           const float* __restrict__ cond, int Cc, int B, int H,
                                                           int W, uint32_t* __restrict__ out) {
  pdl_prologue();
  // Work item = 64 consecutive pixels of one image row.  Phase 1: for every pair of k = (tap, channel) the 64 pixel
  // values are read with COALESCED loads along W (lanes = consecutive pixels) and dropped, packed, into a
  // [pixel][k-pair] smem tile.  Phase 2: the tile is streamed out as full 256-byte rows (coalesced 4-byte words).
  // The k -> (plane, dh, dw) decode is done once per block into a small table; blocks loop over work items.
  __shared__ uint32_t tile[kI2cPix * kI2cPitchW];
  __shared__ int s_chan[128];   // channel index (>= 0: sample plane c; < -1: cond plane -(v+2); -1: zero padding)
  __shared__ int s_off[128];    // (dh + 1) * 3 + (dw + 1) = tap index
  const int Cin = Cs + Cc;
  const int K = 9 * Cin;
  if (threadIdx.x < 128) {
    const int k = threadIdx.x;
    int chan = -1, code = 4;
    if (k < K) {
      const int tap = k / Cin;
      const int c = k - tap * Cin;
      chan = (c < Cs) ? c : -(c - Cs + 2);
      code = tap;
    }
    s_chan[k] = chan;
    s_off[k] = code;
  }
  __syncthreads();
  const int segs = (W + kI2cPix - 1) / kI2cPix;
  const int items = B * H * segs;
  const size_t HWs = (size_t)H * W;
  const int px = threadIdx.x & 63;   // pixel inside the segment
  const int kq = threadIdx.x >> 6;   // 4 k-pair lanes per pixel
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int seg = item % segs;
    const int h = (item / segs) % H;
    const size_t b = item / (segs * H);
    const int w0 = seg * kI2cPix;
    const int w = w0 + px;
    const float* xb = xs + b * Cs * HWs;
    const float* cb = cond + b * Cc * HWs;
#pragma unroll 4
    for (int kp = kq; kp < 64; kp += 4) {
      float val[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * kp + e;
        const int chan = s_chan[k];
        const int code = s_off[k];
        const int dh = (code >= 6 ? 1 : (code >= 3 ? 0 : -1)), dw = code - 3 * (dh + 1) - 1;
        const int hh = h + dh, ww = w + dw;
        float v = 0.f;
        if (chan != -1 && hh >= 0 && hh < H && ww >= 0 && ww < W) {
          const float* src = (chan >= 0) ? xb + (size_t)chan * HWs : cb + (size_t)(-chan - 2) * HWs;
          v = __ldg(src + (size_t)hh * W + ww);
        }
        val[e] = v;
      }
      tile[px * kI2cPitchW + kp] = pack2<DT>(val[0], val[1]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kI2cPix * 64; i += 256) {
      const int p = i >> 6, wd = i & 63;
      if (w0 + p < W) __stcs(out + ((b * H + h) * (size_t)W + w0 + p) * 64 + wd, tile[p * kI2cPitchW + wd]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// direct conv_in (split form): conv_in(cat[x, cond]) = conv(x; W[:, :Cs]) + {conv(cond; W[:, Cs:]) + bias}.  The braces
// depend only on the conditioning images, which are constant over the steps of one sampling loop
// (custom_pipelines.py:158-171 rebuilds the same cat every step), so they are computed ONCE per loop into an fp32
// NHWC "base" tensor (OUT32 mode) and every step only adds the 9*Cs taps of the noisy sample (OUT16 mode: + base,
// round to the 16-bit activation type, GroupNorm quad statistics of the rounded values).  All arithmetic is fp32 with
// fp32 weights - the model input is never rounded to 16 bit.
// Per step algorithmic bytes: (4 (base) + 2 (out)) * 128 * P + 4 * Cs * P per sample.
// Thread = 4 consecutive pixels of a row x 8 output channels; block = 64 pixels x 128 channels; weights in smem.
// ------------------------------------------------------------------------------------------------
constexpr int kCiPx = 4;          // pixels per thread
constexpr int kCiSeg = 16 * kCiPx;  // pixels per block iteration

struct ConvInArgs {
  const float* planes;  // [B, C, H, W] fp32
  int C;
  const float* wT;      // [C*9][128] fp32, k = c*9 + tap
  const void* base;     // [B,H,W,128] fp32 (base_f16 == 0) or fp16 (base_f16 == 1), or nullptr
  int base_f16;
  const float* bias;    // [128] or nullptr (added when base == nullptr)
  float* out32;         // OUT32 mode
  uint16_t* out16;      // OUT16 mode
  double* stats;        // OUT16 mode: [B][32][2]
  int B, H, W;
};

template <int DT, bool OUT32>
__global__ void __launch_bounds__(256, 2) conv_in_direct_kernel(const ConvInArgs a) {
  extern __shared__ __align__(16) float ci_w[];  // [C*9][128]
  __shared__ float s_part[8][64];
  const int K = a.C * 9;
  for (int i = threadIdx.x; i < K * 128; i += 256) ci_w[i] = a.wT[i];   // static weights: staged BEFORE the PDL wait
  __syncthreads();
  pdl_prologue();
  const int oct = threadIdx.x & 15, pg = threadIdx.x >> 4;
  const int b = blockIdx.y;
  const int segs = (a.W + kCiSeg - 1) / kCiSeg;
  const int items = a.H * segs;
  const size_t HW = (size_t)a.H * a.W;
  float bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias8[j] = (a.base == nullptr && a.bias != nullptr) ? a.bias[oct * 8 + j] : 0.f;
  float st[4] = {0.f, 0.f, 0.f, 0.f};  // quad 2*oct: sum, sq; quad 2*oct+1: sum, sq
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int h = item / segs, seg = item - h * segs;
    const int w0 = seg * kCiSeg + pg * kCiPx;
    if (w0 >= a.W) continue;
    float acc[kCiPx][8];
#pragma unroll
    for (int p = 0; p < kCiPx; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = bias8[j];
    // issue the loads of the per-loop term first so that they are in flight during the tap arithmetic
    float4 bv[kCiPx][2];
    if (a.base != nullptr && a.base_f16) {
      // 16-bit per-loop term (fp16 in both operand modes: |values| = O(1), 11 significant bits - below the rounding of
      // the 16-bit activation this kernel stores): (2 + 2) * 128 * P bytes per step instead of (4 + 2) * 128 * P
      const __half* bp = static_cast<const __half*>(a.base) + (((size_t)b * a.H + h) * a.W + w0) * 128 + oct * 8;
#pragma unroll
      for (int p = 0; p < kCiPx; ++p) {
        if (w0 + p < a.W) {
          const uint4 u = __ldcs(reinterpret_cast<const uint4*>(bp + (size_t)p * 128));
          const float2 f0 = unpack2<kF16>(u.x), f1 = unpack2<kF16>(u.y), f2 = unpack2<kF16>(u.z), f3 = unpack2<kF16>(u.w);
          bv[p][0] = make_float4(f0.x, f0.y, f1.x, f1.y);
          bv[p][1] = make_float4(f2.x, f2.y, f3.x, f3.y);
        } else {
          bv[p][0] = bv[p][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else if (a.base != nullptr) {
      const float* bp = static_cast<const float*>(a.base) + (((size_t)b * a.H + h) * a.W + w0) * 128 + oct * 8;
#pragma unroll
      for (int p = 0; p < kCiPx; ++p) {
        if (w0 + p < a.W) {
          bv[p][0] = __ldcs(reinterpret_cast<const float4*>(bp + (size_t)p * 128));
          bv[p][1] = __ldcs(reinterpret_cast<const float4*>(bp + (size_t)p * 128 + 4));
        } else {
          bv[p][0] = bv[p][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    for (int c = 0; c < a.C; ++c) {
      const float* pl = a.planes + ((size_t)b * a.C + c) * HW;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        float xv[kCiPx + 2];
#pragma unroll
        for (int i = 0; i < kCiPx + 2; ++i) {
          const int ww = w0 - 1 + i;
          xv[i] = (hh >= 0 && hh < a.H && ww >= 0 && ww < a.W) ? __ldg(pl + (size_t)hh * a.W + ww) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wa = *reinterpret_cast<const float4*>(ci_w + (c * 9 + kh * 3 + kw) * 128 + oct * 8);
          const float4 wb = *reinterpret_cast<const float4*>(ci_w + (c * 9 + kh * 3 + kw) * 128 + oct * 8 + 4);
#pragma unroll
          for (int p = 0; p < kCiPx; ++p) {
            const float x = xv[p + kw];
            acc[p][0] = fmaf(x, wa.x, acc[p][0]); acc[p][1] = fmaf(x, wa.y, acc[p][1]);
            acc[p][2] = fmaf(x, wa.z, acc[p][2]); acc[p][3] = fmaf(x, wa.w, acc[p][3]);
            acc[p][4] = fmaf(x, wb.x, acc[p][4]); acc[p][5] = fmaf(x, wb.y, acc[p][5]);
            acc[p][6] = fmaf(x, wb.z, acc[p][6]); acc[p][7] = fmaf(x, wb.w, acc[p][7]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kCiPx; ++p) {
      const int w = w0 + p;
      if (w >= a.W) break;
      const size_t pix = ((size_t)b * a.H + h) * a.W + w;
      if (a.base != nullptr) {
        const float4 ba = bv[p][0], bb = bv[p][1];
        acc[p][0] += ba.x; acc[p][1] += ba.y; acc[p][2] += ba.z; acc[p][3] += ba.w;
        acc[p][4] += bb.x; acc[p][5] += bb.y; acc[p][6] += bb.z; acc[p][7] += bb.w;
      }
      if constexpr (OUT32) {
        float4* o = reinterpret_cast<float4*>(a.out32 + pix * 128 + oct * 8);
        o[0] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        o[1] = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
      } else {
        uint4 pk;
        pk.x = pack2<DT>(acc[p][0], acc[p][1]);
        pk.y = pack2<DT>(acc[p][2], acc[p][3]);
        pk.z = pack2<DT>(acc[p][4], acc[p][5]);
        pk.w = pack2<DT>(acc[p][6], acc[p][7]);
        *reinterpret_cast<uint4*>(a.out16 + pix * 128 + oct * 8) = pk;
        const float2 f0 = unpack2<DT>(pk.x), f1 = unpack2<DT>(pk.y), f2 = unpack2<DT>(pk.z), f3 = unpack2<DT>(pk.w);
        st[0] += (f0.x + f0.y) + (f1.x + f1.y);
        st[1] += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
        st[2] += (f2.x + f2.y) + (f3.x + f3.y);
        st[3] += (f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y);
      }
    }
  }
  if constexpr (!OUT32) {
    if (a.stats != nullptr) {
      // lanes l and l+16 of a warp share the octet: combine, park per warp, then a fixed-order sum over the 8 warps
      // (deterministic: no floating-point atomics before the fp64 stage)
#pragma unroll
      for (int i = 0; i < 4; ++i) st[i] += __shfl_xor_sync(0xffffffffu, st[i], 16);
      if ((threadIdx.x & 31) < 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_part[threadIdx.x >> 5][oct * 4 + i] = st[i];
      }
      __syncthreads();
      if (threadIdx.x < 64) {
        // s_part column oct*4 + {0,1,2,3} = quad 2*oct {sum, sq}, quad 2*oct+1 {sum, sq} = stats[b][quad][2] flattened
        double t = 0.0;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) t += (double)s_part[wi][threadIdx.x];
        atomicAdd(a.stats + (size_t)b * 64 + threadIdx.x, t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused GroupNorm + SiLU + conv3x3 (128 -> DC) -> fp32 NCHW, in two kernels without any halo:
//   A  taps:   y[b][d*9+tap][p] = sum_c w[d][tap][c] * silu(gn(x[b][p][c]))   for every pixel p and all 9*DC taps
//              = a streaming GEMM [P x 128] x [128 x 9*DC]: each activation is read ONCE from HBM straight into MMA
//              A-fragment registers (no shared memory), normalised + SiLU'd + rounded there (what the separate
//              GroupNorm pass would have stored), multiplied by register-resident weight fragments (mma.m16n8k16).
//   B  gather: out[b][d][h][w] = bias[d] + sum_tap y[b][d*9+tap][h+dh][w+dw]   (zero outside the image = the conv's
//              zero padding of the normalised activation).
// The earlier one-kernel version re-read every activation 9 times from shared memory (ldmatrix per tap: 0.38 ms of
// shared-memory time) and normalised a 33 % halo.  Algorithmic bytes: (128 * 2 + 2 * 9*DC * 4 + DC * 4) * P per sample.
//
// K permutation: lane (g, t) owns the 16-byte chunks 4q + t (q = 0..3; 8 channels each) of pixels g and g+8, so one
// load instruction of a warp covers 8 pixels x 64 contiguous bytes (whole sectors).  Chunk q feeds the m16n8k16 slabs
// 2q (channels +0..3 as k = 2t, 2t+1, 2t+8, 2t+9) and 2q+1 (channels +4..7); the weight fragments are permuted the
// same way.
// ------------------------------------------------------------------------------------------------
constexpr int kCoThreads = 256;

struct ConvOutArgs {
  const uint16_t* x;  // [B,H,W,128]
  int pitch;
  const double* stats;     // [B][32][2] quad stats (4 ch per group here)
  const float* gamma;
  const float* beta;
  const float* w;          // [DC][9][128] fp32 (tap-major, channel fastest)
  const float* bias;       // [DC]
  float* out;              // [B,DC,H,W]
  float* taps;             // workspace [B][9*DC][H*W] fp32
  int B, H, W;
  float eps;
};

template <int DT>
D3_DEVINL void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (DT == kBF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

template <int DT>
D3_DEVINL uint32_t norm_silu_pair(uint32_t raw, float s0, float h0, float s1, float h1) {
  const float2 f = unpack2<DT>(raw);
  return pack2<DT>(silu_fast(fmaf(f.x, s0, h0)), silu_fast(fmaf(f.y, s1, h1)));
}

template <int DC, int DT>
__global__ void __launch_bounds__(kCoThreads, (DC == 1) ? 2 : 1) conv_out_taps_kernel(const ConvOutArgs a) {
  pdl_prologue();
  constexpr int NTAPS = 9 * DC;
  constexpr int NT = (NTAPS + 7) / 8;  // n-tiles of 8 taps
  __shared__ float s_sc[128], s_sh[128];
  const int b = blockIdx.y;
  const int HW = a.H * a.W;
  if (threadIdx.x < 32) {
    const int grp = threadIdx.x;  // 4 channels per group at C = 128
    const double s = a.stats[((size_t)b * 32 + grp) * 2], ss = a.stats[((size_t)b * 32 + grp) * 2 + 1];
    const double n = (double)HW * 4.0;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = grp * 4 + k;
      const float scv = a.gamma[c] * rstd;
      s_sc[c] = scv;
      s_sh[c] = a.beta[c] - (float)mean * scv;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // weight fragments: slab s, n-tile j: b0 = {w[n][32t+4s], w[n][32t+4s+1]}, b1 = {.. +2, .. +3}, n = 8j + g
  uint32_t wf[8][NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int n = 8 * j + g;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (n < NTAPS) {
        const float* wp = a.w + (size_t)n * 128 + 8 * (4 * (sl >> 1) + t) + 4 * (sl & 1);  // [n][c]
        w0 = wp[0]; w1 = wp[1]; w2 = wp[2]; w3 = wp[3];
      }
      wf[sl][j][0] = pack2<DT>(w0, w1);
      wf[sl][j][1] = pack2<DT>(w2, w3);
    }
  }
  const int n_groups = (HW + 15) / 16;
  const uint16_t* xb = a.x + (size_t)b * HW * a.pitch;
  float* yb = a.taps + (size_t)b * NTAPS * HW;
  // software pipeline: the 8 x 16-byte loads of the warp's NEXT 16-pixel group are issued before the current group's
  // arithmetic; scale/shift are re-read from shared memory per 8-channel vector to leave registers for that.
  const float4* sc4 = reinterpret_cast<const float4*>(s_sc + 8 * t);  // + 32 floats per q
  const float4* sh4 = reinterpret_cast<const float4*>(s_sh + 8 * t);
  const int stride16 = gridDim.x * (kCoThreads / 32);
  int grp16 = blockIdx.x * (kCoThreads / 32) + warp;
  uint4 r0[4], r1[4];
  auto load_group = [&](int gi) {
    const int q0 = gi * 16 + g, q1 = q0 + 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r0[q] = (q0 < HW) ? __ldcs(reinterpret_cast<const uint4*>(xb + (size_t)q0 * a.pitch + 8 * t) + 4 * q)
                        : make_uint4(0u, 0u, 0u, 0u);
      r1[q] = (q1 < HW) ? __ldcs(reinterpret_cast<const uint4*>(xb + (size_t)q1 * a.pitch + 8 * t) + 4 * q)
                        : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (grp16 < n_groups) load_group(grp16);
  for (; grp16 < n_groups; grp16 += stride16) {
    const int p0 = grp16 * 16 + g, p1 = p0 + 8;
    uint4 c0[4], c1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { c0[q] = r0[q]; c1[q] = r1[q]; }
    if (grp16 + stride16 < n_groups) load_group(grp16 + stride16);
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // uint4 q holds channels 8(4q+t) .. +7: slab 2q = words x, y; slab 2q+1 = words z, w
      const float4 sa = sc4[8 * q], sb = sc4[8 * q + 1], ha = sh4[8 * q], hb = sh4[8 * q + 1];
      uint32_t af[4];
      af[0] = norm_silu_pair<DT>(c0[q].x, sa.x, ha.x, sa.y, ha.y);
      af[1] = norm_silu_pair<DT>(c1[q].x, sa.x, ha.x, sa.y, ha.y);
      af[2] = norm_silu_pair<DT>(c0[q].y, sa.z, ha.z, sa.w, ha.w);
      af[3] = norm_silu_pair<DT>(c1[q].y, sa.z, ha.z, sa.w, ha.w);
#pragma unroll
      for (int j = 0; j < NT; ++j) mma16816<DT>(acc[j], af, wf[2 * q][j][0], wf[2 * q][j][1]);
      af[0] = norm_silu_pair<DT>(c0[q].z, sb.x, hb.x, sb.y, hb.y);
      af[1] = norm_silu_pair<DT>(c1[q].z, sb.x, hb.x, sb.y, hb.y);
      af[2] = norm_silu_pair<DT>(c0[q].w, sb.z, hb.z, sb.w, hb.w);
      af[3] = norm_silu_pair<DT>(c1[q].w, sb.z, hb.z, sb.w, hb.w);
#pragma unroll
      for (int j = 0; j < NT; ++j) mma16816<DT>(acc[j], af, wf[2 * q + 1][j][0], wf[2 * q + 1][j][1]);
    }
    // C fragment: rows g / g+8 (pixels), columns 2t, 2t+1 (taps of n-tile j) -> planar tap maps
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 8 * j + 2 * t + e;
        if (n < NTAPS) {
          if (p0 < HW) yb[(size_t)n * HW + p0] = acc[j][e];
          if (p1 < HW) yb[(size_t)n * HW + p1] = acc[j][2 + e];
        }
      }
    }
  }
}

// FUSE: the gathered value IS the model output of the step; apply the DDIM update to it right here (sample in, x_{t-1}
// and x0 out: scheduler_ddim.py:423-459 with eta = 0) instead of writing it for a separate step kernel to read back.
struct GatherDdim {
  const float* sample;   // x_t, same [B,DC,H,W] layout as the model output
  d3_ddim_coeffs c;
  float* prev;           // x_{t-1}
  float* orig;           // predicted x0
};

template <int DC, bool FUSE>
__global__ void __launch_bounds__(256) conv_out_gather_kernel(const float* __restrict__ taps,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int B, int H, int W, const GatherDdim f) {
  pdl_prologue();
  const int HW = H * W;
  const size_t total = (size_t)B * DC * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int d = (int)((i / HW) % DC);
    const size_t b = i / ((size_t)HW * DC);
    const int h = p / W, w = p - h * W;
    const float* yb = taps + (b * 9 * DC + (size_t)d * 9) * HW;
    float acc = bias[d];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc += __ldg(yb + (size_t)tap * HW + (size_t)hh * W + ww);
    }
    if constexpr (FUSE) {
      const float s = f.sample[i];
      float x0, eps;
      split_x0_eps(f.c.prediction_type, acc, s, f.c.sqrt_alpha_t, f.c.sqrt_beta_t, x0, eps);
      x0 = clip_x0(x0, f.c.clip_mode, f.c.clip_range);
      if (f.c.use_clipped_model_output) eps = __fdiv_rn(__fsub_rn(s, __fmul_rn(f.c.sqrt_alpha_t, x0)), f.c.sqrt_beta_t);
      f.orig[i] = x0;
      f.prev[i] = __fadd_rn(__fmul_rn(f.c.sqrt_alpha_prev, x0), __fmul_rn(f.c.dir_coeff, eps));
      if (out != nullptr) out[i] = acc;
    } else {
      out[i] = acc;
    }
  }
}

}  // namespace
}  // namespace d3

extern "C" int d3_im2col_input(const float* sample, int Cs, const float* cond, int Cc, int B, int H, int W,
                               void* out, int dtype, void* stream) {
  using namespace d3;
  D3_REQUIRE(sample && out && (Cc == 0 || cond), "im2col: null pointer");
  D3_REQUIRE(9 * (Cs + Cc) <= 128, "im2col: 9*Cin = %d exceeds the 128-wide K slab", 9 * (Cs + Cc));
  if (dtype == D3_DTYPE_F32) return im2col_f32(sample, Cs, cond, Cc, B, H, W, out, static_cast<cudaStream_t>(stream));
  int blocks = B * H * ((W + kI2cPix - 1) / kI2cPix);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (dtype == D3_DTYPE_F16)
    D3_CHECK_CUDA(launch_pdl(im2col_input_kernel<kF16>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), sample, Cs, cond, Cc, B, H, W, static_cast<uint32_t*>(out)));
  else
    D3_CHECK_CUDA(launch_pdl(im2col_input_kernel<kBF16>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), sample, Cs, cond, Cc, B, H, W, static_cast<uint32_t*>(out)));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" size_t d3_conv_out_workspace_bytes(int B, int H, int W, int dc) {
  return (size_t)B * 9 * dc * H * W * sizeof(float);
}

namespace d3 {
namespace {
int conv_out_impl(const void* x, int pitch, const double* stats, const float* gamma, const float* beta, const float* w,
                  const float* bias, int B, int H, int W, int dc, float eps, float* out, void* workspace, int dtype,
                  const GatherDdim* fuse, void* stream) {
  D3_REQUIRE(x && stats && gamma && beta && w && bias && (out || fuse), "conv_out: null pointer");
  D3_REQUIRE(dc == 1 || dc == 3, "conv_out: depth_channels must be 1 or 3 (got %d)", dc);
  if (dtype == D3_DTYPE_F32) {
    D3_REQUIRE(fuse == nullptr, "conv_out: the fused DDIM update exists for the 16-bit paths only");
    return conv_out_f32(x, pitch, stats, gamma, beta, w, bias, B, H, W, dc, eps, out, static_cast<cudaStream_t>(stream));
  }
  D3_REQUIRE(workspace != nullptr, "conv_out: workspace of d3_conv_out_workspace_bytes() bytes required");
  D3_REQUIRE(pitch % 8 == 0, "conv_out: pitch must be a multiple of 8");
  ConvOutArgs a{static_cast<const uint16_t*>(x), pitch, stats, gamma, beta, w, bias, out,
                static_cast<float*>(workspace), B, H, W, eps};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_groups = (H * W + 15) / 16;
  const int per_block = kCoThreads / 32;
  int gx = (num_sms() * (dc == 1 ? 2 : 1) * 2 + B - 1) / B;   // ~2 waves of resident blocks over the whole batch
  if (gx > (n_groups + per_block - 1) / per_block) gx = (n_groups + per_block - 1) / per_block;
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  if (dc == 1 && dtype == D3_DTYPE_BF16) D3_CHECK_CUDA(launch_pdl(conv_out_taps_kernel<1, kBF16>, dim3(grid), dim3(kCoThreads), 0, st, a));
  else if (dc == 1) D3_CHECK_CUDA(launch_pdl(conv_out_taps_kernel<1, kF16>, dim3(grid), dim3(kCoThreads), 0, st, a));
  else if (dtype == D3_DTYPE_BF16) D3_CHECK_CUDA(launch_pdl(conv_out_taps_kernel<3, kBF16>, dim3(grid), dim3(kCoThreads), 0, st, a));
  else D3_CHECK_CUDA(launch_pdl(conv_out_taps_kernel<3, kF16>, dim3(grid), dim3(kCoThreads), 0, st, a));
  D3_LAUNCH_CHECK();
  const size_t total = (size_t)B * dc * H * W;
  int gb = (int)((total + 255) / 256);
  if (gb > num_sms() * 16) gb = num_sms() * 16;
  GatherDdim f{};
  if (fuse) f = *fuse;
  if (fuse && dc == 1) D3_CHECK_CUDA(launch_pdl(conv_out_gather_kernel<1, true>, dim3(gb), dim3(256), 0, st, a.taps, bias, out, B, H, W, f));
  else if (fuse) D3_CHECK_CUDA(launch_pdl(conv_out_gather_kernel<3, true>, dim3(gb), dim3(256), 0, st, a.taps, bias, out, B, H, W, f));
  else if (dc == 1) D3_CHECK_CUDA(launch_pdl(conv_out_gather_kernel<1, false>, dim3(gb), dim3(256), 0, st, a.taps, bias, out, B, H, W, f));
  else D3_CHECK_CUDA(launch_pdl(conv_out_gather_kernel<3, false>, dim3(gb), dim3(256), 0, st, a.taps, bias, out, B, H, W, f));
  D3_LAUNCH_CHECK();
  return D3_OK;
}
}  // namespace
}  // namespace d3

extern "C" int d3_conv_out_fused(const void* x, int pitch, const double* stats, const float* gamma,
                                 const float* beta, const float* w, const float* bias, int B, int H, int W, int dc,
                                 float eps, float* out, void* workspace, int dtype, void* stream) {
  D3_REQUIRE(out != nullptr, "conv_out: null output");
  return d3::conv_out_impl(x, pitch, stats, gamma, beta, w, bias, B, H, W, dc, eps, out, workspace, dtype, nullptr, stream);
}

extern "C" int d3_conv_out_fused_ddim(const void* x, int pitch, const double* stats, const float* gamma,
                                      const float* beta, const float* w, const float* bias, int B, int H, int W, int dc,
                                      float eps, float* model_out, void* workspace, int dtype, const float* sample,
                                      const d3_ddim_coeffs* c, float* prev_sample, float* pred_original,
                                      void* stream) {
  D3_REQUIRE(sample && c && prev_sample && pred_original, "conv_out_ddim: null pointer");
  D3_REQUIRE(c->std_dev == 0.0f, "conv_out_ddim: the fused update is the eta = 0 step (no variance noise)");
  d3::GatherDdim f{sample, *c, prev_sample, pred_original};
  return d3::conv_out_impl(x, pitch, stats, gamma, beta, w, bias, B, H, W, dc, eps, model_out, workspace, dtype, &f,
                           stream);
}

extern "C" int d3_conv_in_direct(const float* planes, int C, const float* wT, const void* base, int base_dtype,
                                 const float* bias, int B, int H, int W, float* out32, void* out16, int dtype,
                                 double* stats, void* stream) {
  using namespace d3;
  D3_REQUIRE(planes && wT, "conv_in_direct: null pointer");
  D3_REQUIRE(C >= 1 && C <= 12, "conv_in_direct: C must be in [1, 12] (got %d)", C);
  D3_REQUIRE((out32 != nullptr) != (out16 != nullptr), "conv_in_direct: exactly one of out32 / out16 must be given");
  D3_REQUIRE(out32 || dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "conv_in_direct: bad dtype %d", dtype);
  D3_REQUIRE(base || bias, "conv_in_direct: base or bias required");
  D3_REQUIRE(!base || base_dtype == D3_DTYPE_F32 || base_dtype == D3_DTYPE_F16,
             "conv_in_direct: base must be fp32 or fp16 (got dtype %d)", base_dtype);
  ConvInArgs a{planes, C, wT, base, base_dtype == D3_DTYPE_F16 ? 1 : 0, bias, out32, static_cast<uint16_t*>(out16),
               stats, B, H, W};
  const int items = H * ((W + kCiSeg - 1) / kCiSeg);
  int gx = (num_sms() * 4 + B - 1) / B;
  if (gx > items) gx = items;
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  const size_t smem = (size_t)C * 9 * 128 * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t mx = 12 * 9 * 128 * 4;
  D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_in_direct_kernel<kBF16, true>), mx));
  D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_in_direct_kernel<kBF16, false>), mx));
  D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(conv_in_direct_kernel<kF16, false>), mx));
  if (out32) D3_CHECK_CUDA(launch_pdl(conv_in_direct_kernel<kBF16, true>, dim3(grid), dim3(256), smem, st, a));
  else if (dtype == D3_DTYPE_BF16) D3_CHECK_CUDA(launch_pdl(conv_in_direct_kernel<kBF16, false>, dim3(grid), dim3(256), smem, st, a));
  else D3_CHECK_CUDA(launch_pdl(conv_in_direct_kernel<kF16, false>, dim3(grid), dim3(256), smem, st, a));
  D3_LAUNCH_CHECK();
  return D3_OK;
}
