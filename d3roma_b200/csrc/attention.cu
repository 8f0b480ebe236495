// Spatial self-attention core of the UNet attention blocks: 64 heads x head_dim 8,
// N = H*W tokens (960 at 24x40, 240 at 12x20, 2040 at 34x60), softmax(q k^T / sqrt(8)) v.
// Replaces aten::scaled_dot_product_attention as invoked by diffusers' Attention processor (blocks built at
// /root/reference/train.py:788-810).  The GroupNorm, the fused QKV projection and the output projection + residual
// run on the tcgen05 igemm kernel; this kernel is the softmax(QK^T)V part.
//
// Round-1 implementation: CUDA-core flash-style kernel, one query per thread, the whole K/V of one (sample, head)
// staged in shared memory as fp32 (15-65 KB), online softmax in the exp2 domain.  head_dim = 8 makes the tensor-core
// formulation awkward (K padded 8 -> 16, PV N = 8) and the op is exp-bound; a tcgen05 version is future work and is
// tracked in DESIGN.md.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kAttnThreads = 256;
constexpr int kHeadDim = 8;

D3_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DT>
__global__ void __launch_bounds__(kAttnThreads) attention_hd8_kernel(const uint16_t* __restrict__ qkv, int pitch,
                                                                     int C, int N, uint16_t* __restrict__ out,
                                                                     int opitch) {
  extern __shared__ float4 s_kv[];  // [N][2] float4 for K, then [N][2] float4 for V
  float4* s_k = s_kv;
  float4* s_v = s_kv + 2 * N;
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = qkv + (size_t)b * N * pitch + head * kHeadDim;
  for (int j = threadIdx.x; j < N; j += kAttnThreads) {
    const uint4 kr = __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * pitch + C));
    const uint4 vr = __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * pitch + 2 * C));
    float2 a = unpack2<DT>(kr.x), c = unpack2<DT>(kr.y), d = unpack2<DT>(kr.z), e = unpack2<DT>(kr.w);
    s_k[2 * j] = make_float4(a.x, a.y, c.x, c.y);
    s_k[2 * j + 1] = make_float4(d.x, d.y, e.x, e.y);
    a = unpack2<DT>(vr.x); c = unpack2<DT>(vr.y); d = unpack2<DT>(vr.z); e = unpack2<DT>(vr.w);
    s_v[2 * j] = make_float4(a.x, a.y, c.x, c.y);
    s_v[2 * j + 1] = make_float4(d.x, d.y, e.x, e.y);
  }
  __syncthreads();
  const int qi = blockIdx.x * kAttnThreads + threadIdx.x;
  if (qi >= N) return;
  float q[8];
  {
    const uint4 qr = __ldg(reinterpret_cast<const uint4*>(base + (size_t)qi * pitch));
    const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)
    float2 a = unpack2<DT>(qr.x), c = unpack2<DT>(qr.y), d = unpack2<DT>(qr.z), e = unpack2<DT>(qr.w);
    q[0] = a.x * sc; q[1] = a.y * sc; q[2] = c.x * sc; q[3] = c.y * sc;
    q[4] = d.x * sc; q[5] = d.y * sc; q[6] = e.x * sc; q[7] = e.y * sc;
  }
  float m = -INFINITY, l = 0.f;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int KB = 8;
  for (int j0 = 0; j0 < N; j0 += KB) {
    float s[KB];
    float mx = m;
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = min(j0 + u, N - 1);
      const float4 k0 = s_k[2 * j], k1 = s_k[2 * j + 1];
      float acc = q[0] * k0.x;
      acc = fmaf(q[1], k0.y, acc); acc = fmaf(q[2], k0.z, acc); acc = fmaf(q[3], k0.w, acc);
      acc = fmaf(q[4], k1.x, acc); acc = fmaf(q[5], k1.y, acc); acc = fmaf(q[6], k1.z, acc);
      acc = fmaf(q[7], k1.w, acc);
      s[u] = (j0 + u < N) ? acc : -INFINITY;
      mx = fmaxf(mx, s[u]);
    }
    const float corr = fast_exp2(m - mx);  // exp2(-inf) = 0 on the first block
    m = mx;
    l *= corr;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] *= corr;
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = min(j0 + u, N - 1);
      const float p = fast_exp2(s[u] - m);
      l += p;
      const float4 v0 = s_v[2 * j], v1 = s_v[2 * j + 1];
      o[0] = fmaf(p, v0.x, o[0]); o[1] = fmaf(p, v0.y, o[1]); o[2] = fmaf(p, v0.z, o[2]); o[3] = fmaf(p, v0.w, o[3]);
      o[4] = fmaf(p, v1.x, o[4]); o[5] = fmaf(p, v1.y, o[5]); o[6] = fmaf(p, v1.z, o[6]); o[7] = fmaf(p, v1.w, o[7]);
    }
  }
  const float inv = 1.0f / l;
  uint4 r;
  r.x = pack2<DT>(o[0] * inv, o[1] * inv);
  r.y = pack2<DT>(o[2] * inv, o[3] * inv);
  r.z = pack2<DT>(o[4] * inv, o[5] * inv);
  r.w = pack2<DT>(o[6] * inv, o[7] * inv);
  *reinterpret_cast<uint4*>(out + ((size_t)b * N + qi) * opitch + head * kHeadDim) = r;
}

}  // namespace
}  // namespace d3

extern "C" int d3_attention_hd8(const void* qkv, int pitch, int B, int N, int C, void* out, int opitch,
                                int dtype, void* stream) {
  using namespace d3;
  D3_REQUIRE(qkv && out && C % 8 == 0 && pitch >= 3 * C && pitch % 8 == 0 && opitch % 8 == 0,
             "attention: bad arguments");
  const size_t smem = (size_t)N * 64;
  D3_REQUIRE(smem <= 220 * 1024, "attention: N=%d tokens exceed the shared-memory K/V budget", N);
  static size_t configured = 0;
  if (smem > configured) {
    D3_CHECK_CUDA(cudaFuncSetAttribute(attention_hd8_kernel<kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    D3_CHECK_CUDA(cudaFuncSetAttribute(attention_hd8_kernel<kF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    configured = smem;
  }
  dim3 grid(ceil_div(N, kAttnThreads), C / kHeadDim, B);
  if (dtype == D3_DTYPE_F16)
    attention_hd8_kernel<kF16><<<grid, kAttnThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(qkv), pitch, C, N, static_cast<uint16_t*>(out), opitch);
  else
    attention_hd8_kernel<kBF16><<<grid, kAttnThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(qkv), pitch, C, N, static_cast<uint16_t*>(out), opitch);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
