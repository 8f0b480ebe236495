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
  // 16 lanes per pixel; lane j builds k = 8j .. 8j+7 of the 128-wide row.  The (tap, channel) of each of its 8
  // elements does not depend on the pixel, so it is decoded once: plane base pointer offset + (dh, dw).
  const int Cin = Cs + Cc;
  const int K = 9 * Cin;
  const size_t npix = (size_t)B * H * W;
  const size_t HWs = (size_t)H * W;
  const int j = threadIdx.x & 15;
  int dh[8], dw[8], ch[8];  // ch < 0: zero padding column; ch < Cs: sample plane; else cond plane (ch - Cs)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = 8 * j + e;
    if (k < K) {
      const int tap = k / Cin;
      ch[e] = k - tap * Cin;
      dh[e] = tap / 3 - 1;
      dw[e] = tap % 3 - 1;
    } else {
      ch[e] = -1;
      dh[e] = dw[e] = 0;
    }
  }
  for (size_t p = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; p < npix;
       p += ((size_t)gridDim.x * blockDim.x) >> 4) {
    const int w = (int)(p % W);
    const size_t t = p / W;
    const int h = (int)(t % H);
    const size_t b = t / H;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int hh = h + dh[e], ww = w + dw[e];
      float val = 0.f;
      if (ch[e] >= 0 && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float* src = (ch[e] < Cs) ? xs + (b * Cs + ch[e]) * HWs : cond + (b * Cc + (ch[e] - Cs)) * HWs;
        val = __ldg(src + (size_t)hh * W + ww);
      }
      v[e] = val;
    }
    uint4 o;
    o.x = pack2<DT>(v[0], v[1]);
    o.y = pack2<DT>(v[2], v[3]);
    o.z = pack2<DT>(v[4], v[5]);
    o.w = pack2<DT>(v[6], v[7]);
    __stcs(reinterpret_cast<uint4*>(out + p * 128 + 8 * j), o);
  }
}

// ------------------------------------------------------------------------------------------------
// fused GroupNorm + SiLU + conv3x3 (128 -> DC) -> fp32 NCHW
// ------------------------------------------------------------------------------------------------
constexpr int kCoTY = 8, kCoTX = 32;                 // output tile: 256 pixels, one per thread
constexpr int kCoRY = kCoTY + 2, kCoRX = kCoTX + 2;  // input region with halo
constexpr int kCoPitch = 128 * 2 + 16;               // bytes per pixel row in smem (+16 B pad: conflict-free 16 B reads)
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
  int B, H, W;
  float eps;
};

template <int DC>
constexpr size_t conv_out_smem() {
  return (size_t)kCoRY * kCoRX * kCoPitch + (size_t)DC * 9 * 128 * 4 + 2 * 128 * 4;
}

template <int DC, int DT>
__global__ void __launch_bounds__(kCoThreads) conv_out_kernel(const ConvOutArgs a) {
  // Stage 1: the (8+2) x (32+2) x 128 input region is read ONCE from global (coalesced 16 B vectors), normalised,
  // SiLU'd, rounded to the 16-bit activation type (exactly what the separate GroupNorm pass would have stored) and
  // parked in shared memory; out-of-image pixels become the conv's zero padding.
  // Stage 2: one output pixel per thread: 9 taps x 128 channels from smem against broadcast fp32 weights.
  extern __shared__ __align__(16) uint8_t co_smem[];
  uint8_t* s_in = co_smem;
  float* s_w = reinterpret_cast<float*>(co_smem + (size_t)kCoRY * kCoRX * kCoPitch);
  float* s_sc = s_w + DC * 9 * 128;
  float* s_sh = s_sc + 128;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * kCoTY, x0 = blockIdx.x * kCoTX;
  for (int i = threadIdx.x; i < DC * 9 * 128; i += kCoThreads) s_w[i] = a.w[i];
  if (threadIdx.x < 128) {
    const int c = threadIdx.x, g = c >> 2;
    const double s = a.stats[((size_t)b * 32 + g) * 2], ss = a.stats[((size_t)b * 32 + g) * 2 + 1];
    const double n = (double)a.H * a.W * 4.0;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * rstd;
    s_sc[c] = sc;
    s_sh[c] = a.beta[c] - (float)mean * sc;
  }
  __syncthreads();
  {
    const int vec = threadIdx.x & 15;  // 16 vectors of 8 channels per pixel
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = s_sc[vec * 8 + k];
      sh[k] = s_sh[vec * 8 + k];
    }
    for (int pi = threadIdx.x >> 4; pi < kCoRY * kCoRX; pi += kCoThreads / 16) {
      const int ry = pi / kCoRX, rx = pi - ry * kCoRX;
      const int y = y0 - 1 + ry, x = x0 - 1 + rx;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(a.x + (((size_t)b * a.H + y) * a.W + x) * a.pitch +
                                                               vec * 8));
        const float2 f0 = unpack2<DT>(raw.x), f1 = unpack2<DT>(raw.y);
        const float2 f2 = unpack2<DT>(raw.z), f3 = unpack2<DT>(raw.w);
        o.x = pack2<DT>(silu_fast(fmaf(f0.x, sc[0], sh[0])), silu_fast(fmaf(f0.y, sc[1], sh[1])));
        o.y = pack2<DT>(silu_fast(fmaf(f1.x, sc[2], sh[2])), silu_fast(fmaf(f1.y, sc[3], sh[3])));
        o.z = pack2<DT>(silu_fast(fmaf(f2.x, sc[4], sh[4])), silu_fast(fmaf(f2.y, sc[5], sh[5])));
        o.w = pack2<DT>(silu_fast(fmaf(f3.x, sc[6], sh[6])), silu_fast(fmaf(f3.y, sc[7], sh[7])));
      }
      *reinterpret_cast<uint4*>(s_in + (size_t)pi * kCoPitch + vec * 16) = o;
    }
  }
  __syncthreads();
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  float acc[DC];
#pragma unroll
  for (int d = 0; d < DC; ++d) acc[d] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint8_t* row = s_in + (size_t)((ty + tap / 3) * kCoRX + tx + tap % 3) * kCoPitch;
#pragma unroll 4
    for (int v = 0; v < 16; ++v) {
      const uint4 raw = *reinterpret_cast<const uint4*>(row + v * 16);
      float f[8];
      float2 t2;
      t2 = unpack2<DT>(raw.x); f[0] = t2.x; f[1] = t2.y;
      t2 = unpack2<DT>(raw.y); f[2] = t2.x; f[3] = t2.y;
      t2 = unpack2<DT>(raw.z); f[4] = t2.x; f[5] = t2.y;
      t2 = unpack2<DT>(raw.w); f[6] = t2.x; f[7] = t2.y;
#pragma unroll
      for (int d = 0; d < DC; ++d) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (d * 9 + tap) * 128 + v * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + (d * 9 + tap) * 128 + v * 8 + 4);
        acc[d] = fmaf(f[0], w0.x, acc[d]); acc[d] = fmaf(f[1], w0.y, acc[d]);
        acc[d] = fmaf(f[2], w0.z, acc[d]); acc[d] = fmaf(f[3], w0.w, acc[d]);
        acc[d] = fmaf(f[4], w1.x, acc[d]); acc[d] = fmaf(f[5], w1.y, acc[d]);
        acc[d] = fmaf(f[6], w1.z, acc[d]); acc[d] = fmaf(f[7], w1.w, acc[d]);
      }
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  if (y < a.H && x < a.W) {
#pragma unroll
    for (int d = 0; d < DC; ++d) a.out[(((size_t)b * DC + d) * a.H + y) * a.W + x] = acc[d] + a.bias[d];
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
  ConvOutArgs a{static_cast<const uint16_t*>(x), pitch, stats, gamma, beta, w, bias, out, B, H, W, eps};
  D3_REQUIRE(pitch % 8 == 0, "conv_out: pitch must be a multiple of 8");
  dim3 grid(ceil_div(W, kCoTX), ceil_div(H, kCoTY), B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    D3_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel<1, kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)conv_out_smem<1>()));
    D3_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel<1, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)conv_out_smem<1>()));
    D3_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel<3, kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)conv_out_smem<3>()));
    D3_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel<3, kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)conv_out_smem<3>()));
    attr_set = true;
  }
  if (dc == 1 && dtype == D3_DTYPE_BF16) conv_out_kernel<1, kBF16><<<grid, kCoThreads, conv_out_smem<1>(), st>>>(a);
  else if (dc == 1) conv_out_kernel<1, kF16><<<grid, kCoThreads, conv_out_smem<1>(), st>>>(a);
  else if (dtype == D3_DTYPE_BF16) conv_out_kernel<3, kBF16><<<grid, kCoThreads, conv_out_smem<3>(), st>>>(a);
  else conv_out_kernel<3, kF16><<<grid, kCoThreads, conv_out_smem<3>(), st>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
