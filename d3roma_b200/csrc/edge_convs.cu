// The two "edge" convolutions of the UNet, whose GEMM shapes are degenerate:
//
//  conv_in  (Cin = 5/8/12 -> 128, K = 9*Cin <= 126): the fp32 NCHW model input (noisy sample + conditioning,
//           custom_pipelines.py:158-171) is gathered straight into a bf16 im2col matrix [B,H,W,128]
//           (k = tap*Cin + c, zero-padded to 128) by d3_im2col_input; the tcgen05 igemm then runs it as a 1x1
//           conv with K = 128.  This fuses torch.cat + dtype cast + NCHW->NHWC + zero padding into one pass.
//
//  conv_out (128 -> dc, N = 1 or 3): GroupNorm + SiLU + 3x3 conv fused in one CUDA-core kernel that reads the
//           128-channel activation exactly once per 4-row strip (rolling 3x3 window in registers, one warp per strip,
//           lanes split the channels) and writes the fp32 NCHW model output the schedulers consume.
//           Algorithmic bytes: (128 + dc*2) * P * 2 B per sample.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

// ------------------------------------------------------------------------------------------------
// im2col gather for conv_in
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) im2col_input_kernel(const float* __restrict__ xs, int Cs,
                                                           const float* __restrict__ cond, int Cc, int B, int H,
                                                           int W, uint16_t* __restrict__ out) {
  // 16 lanes per pixel; lane j builds k = 8j .. 8j+7 of the 128-wide row
  const int Cin = Cs + Cc;
  const int K = 9 * Cin;
  const size_t npix = (size_t)B * H * W;
  const int j = threadIdx.x & 15;
  for (size_t p = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; p < npix;
       p += ((size_t)gridDim.x * blockDim.x) >> 4) {
    const int w = (int)(p % W);
    const size_t t = p / W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 8 * j + e;
      float val = 0.f;
      if (k < K) {
        const int tap = k / Cin;
        const int c = k - tap * Cin;
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          val = (c < Cs) ? __ldg(xs + (((size_t)b * Cs + c) * H + hh) * W + ww)
                         : __ldg(cond + (((size_t)b * Cc + (c - Cs)) * H + hh) * W + ww);
        }
      }
      v[e] = val;
    }
    uint4 o;
    o.x = pack2<DT>(v[0], v[1]);
    o.y = pack2<DT>(v[2], v[3]);
    o.z = pack2<DT>(v[4], v[5]);
    o.w = pack2<DT>(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + p * 128 + 8 * j) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// fused GroupNorm + SiLU + conv3x3 (128 -> DC) -> fp32 NCHW
// ------------------------------------------------------------------------------------------------
constexpr int kRows = 4;  // output rows per warp strip

struct ConvOutArgs {
  const uint16_t* x;  // [B,H,W,128]
  int pitch;
  const double* stats;     // [B][32][2] quad stats (4 ch per group here)
  const float* gamma;
  const float* beta;
  const float* w;          // [DC][9][128] fp32 (tap-major, channel fastest)
  const float* bias;       // [DC]
  float* out;              // [B,DC,H,W]
  int B, H, W;
  float eps;
};

template <int DC, int DT>
__global__ void __launch_bounds__(128) conv_out_kernel(const ConvOutArgs a) {
  // one warp per (b, 4-row strip, column segment); lane l owns channels 4l..4l+3 (== GroupNorm group l)
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int strips = ceil_div(a.H, kRows);
  constexpr int kSeg = 64;  // columns per warp task
  const int segs = (a.W + kSeg - 1) / kSeg;
  const int ntasks = a.B * strips * segs;
  if (warp_global >= ntasks) return;
  const int seg = warp_global % segs;
  const int strip = (warp_global / segs) % strips;
  const int b = warp_global / (segs * strips);
  const int r0 = strip * kRows;
  const int c_begin = seg * kSeg;
  const int c_end = min(a.W, c_begin + kSeg);

  // per-lane GroupNorm scale / shift for its 4 channels (group == lane: 128 channels / 32 groups = 4)
  float sc[4], sh[4];
  {
    const double s = a.stats[((size_t)b * 32 + lane) * 2], ss = a.stats[((size_t)b * 32 + lane) * 2 + 1];
    const double n = (double)a.H * a.W * 4.0;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sc[k] = a.gamma[4 * lane + k] * rstd;
      sh[k] = a.beta[4 * lane + k] - (float)mean * sc[k];
    }
  }
  // weights for this lane's channels: [DC][9][4]
  float wr[DC][9][4];
#pragma unroll
  for (int d = 0; d < DC; ++d)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 wv = *reinterpret_cast<const float4*>(a.w + ((size_t)d * 9 + t) * 128 + 4 * lane);
      wr[d][t][0] = wv.x; wr[d][t][1] = wv.y; wr[d][t][2] = wv.z; wr[d][t][3] = wv.w;
    }

  // rolling window: win[row 0..kRows+1][col 0..2][4 ch]
  float win[kRows + 2][3][4];
  auto load_col = [&](int col, int slot) {
#pragma unroll
    for (int r = 0; r < kRows + 2; ++r) {
      const int hh = r0 + r - 1;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (hh >= 0 && hh < a.H && col >= 0 && col < a.W) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(
            a.x + (((size_t)b * a.H + hh) * a.W + col) * a.pitch + 4 * lane));
        const float2 f0 = unpack2<DT>(raw.x), f1 = unpack2<DT>(raw.y);
        v0 = silu_f(fmaf(f0.x, sc[0], sh[0]));
        v1 = silu_f(fmaf(f0.y, sc[1], sh[1]));
        v2 = silu_f(fmaf(f1.x, sc[2], sh[2]));
        v3 = silu_f(fmaf(f1.y, sc[3], sh[3]));
      }
      win[r][slot][0] = v0; win[r][slot][1] = v1; win[r][slot][2] = v2; win[r][slot][3] = v3;
    }
  };

  float keep[DC][kRows];  // lane (col % 32) keeps the result of column col until a coalesced flush
#pragma unroll
  for (int d = 0; d < DC; ++d)
#pragma unroll
    for (int r = 0; r < kRows; ++r) keep[d][r] = 0.f;

  // prologue: columns c_begin-1 and c_begin in slots 0, 1 (slot order rotates through registers below)
  load_col(c_begin - 1, 0);
  load_col(c_begin, 1);
  for (int col = c_begin; col < c_end; ++col) {
    load_col(col + 1, 2);
#pragma unroll
    for (int d = 0; d < DC; ++d) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fmaf(win[r + kh][kw][k], wr[d][kh * 3 + kw][k], acc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == ((col - c_begin) & 31)) keep[d][r] = acc;
      }
    }
    // shift the window left by one column
#pragma unroll
    for (int r = 0; r < kRows + 2; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        win[r][0][k] = win[r][1][k];
        win[r][1][k] = win[r][2][k];
      }
    const int done = col - c_begin + 1;
    if ((done & 31) == 0 || col == c_end - 1) {
      const int base = c_begin + ((done - 1) & ~31);
      const int wcol = base + lane;
      if (wcol <= col) {
#pragma unroll
        for (int d = 0; d < DC; ++d)
#pragma unroll
          for (int r = 0; r < kRows; ++r) {
            const int hh = r0 + r;
            if (hh < a.H) a.out[(((size_t)b * DC + d) * a.H + hh) * a.W + wcol] = keep[d][r] + a.bias[d];
          }
      }
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
  const size_t threads = (size_t)B * H * W * 16;
  int blocks = (int)((threads + 255) / 256);
  const int cap = num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == D3_DTYPE_F16)
    im2col_input_kernel<kF16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sample, Cs, cond, Cc, B, H, W, static_cast<uint16_t*>(out));
  else
    im2col_input_kernel<kBF16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sample, Cs, cond, Cc, B, H, W, static_cast<uint16_t*>(out));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_conv_out_fused(const void* x, int pitch, const double* stats, const float* gamma,
                                 const float* beta, const float* w, const float* bias, int B, int H, int W, int dc,
                                 float eps, float* out, int dtype, void* stream) {
  using namespace d3;
  D3_REQUIRE(x && stats && gamma && beta && w && bias && out, "conv_out: null pointer");
  D3_REQUIRE(dc == 1 || dc == 3, "conv_out: depth_channels must be 1 or 3 (got %d)", dc);
  D3_REQUIRE(pitch % 4 == 0, "conv_out: pitch must be a multiple of 4");
  ConvOutArgs a{static_cast<const uint16_t*>(x), pitch, stats, gamma, beta, w, bias, out, B, H, W, eps};
  const int ntasks = B * ceil_div(H, kRows) * ceil_div(W, 64);
  const int blocks = ceil_div(ntasks, 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dc == 1 && dtype == D3_DTYPE_BF16) conv_out_kernel<1, kBF16><<<blocks, 128, 0, st>>>(a);
  else if (dc == 1) conv_out_kernel<1, kF16><<<blocks, 128, 0, st>>>(a);
  else if (dtype == D3_DTYPE_BF16) conv_out_kernel<3, kBF16><<<blocks, 128, 0, st>>>(a);
  else conv_out_kernel<3, kF16><<<blocks, 128, 0, st>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
