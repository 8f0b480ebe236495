// Backward of the spatial self-attention core (64 heads x head_dim 8; forward: attention.cu):
// given q, k, v (the fused QKV projection output), the forward output o and the gradient d_o, produce dq, dk, dv in
// the layout of the QKV tensor.  Replaces the backward of aten::scaled_dot_product_attention in the reference's
// training step (/root/reference/train.py:455 `accelerator.backward(loss)`).
//
// Two passes, no atomics, probabilities recomputed (flash-attention style, fp32 CUDA-core arithmetic - the op is
// 0.5 % of the UNet's FLOPs):
//   attn_bwd_dq_kernel   one thread per QUERY; K and V of the (sample, head) in shared memory.  Row maximum and sum
//                        (exp2 domain), D_i = d_o_i . o_i, then dq_i = scale * sum_j dS_ij k_j with
//                        dS_ij = p_ij (d_o_i . v_j - D_i).  Row statistics are parked in a small fp32 workspace.
//   attn_bwd_dkv_kernel  one thread per KEY; q, d_o and the row statistics of every query of the (sample, head) in shared
//                        memory.  dv_j = sum_i p_ij d_o_i, dk_j = scale * sum_i dS_ij q_i.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kAbThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

D3_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DT>
D3_DEVINL void load8(const uint16_t* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 a = unpack2<DT>(v.x), b = unpack2<DT>(v.y), c = unpack2<DT>(v.z), d = unpack2<DT>(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <int DT>
D3_DEVINL void store8(uint16_t* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack2<DT>(f[0], f[1]);
  o.y = pack2<DT>(f[2], f[3]);
  o.z = pack2<DT>(f[4], f[5]);
  o.w = pack2<DT>(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

struct AttnBwdArgs {
  const uint16_t* qkv;   // [B, N, 3C]: q | k | v
  int pitch;
  const uint16_t* o;     // [B, N, C]
  int opitch;
  const uint16_t* d_o;   // [B, N, C]
  int dopitch;
  uint16_t* dqkv;        // [B, N, 3C]
  int dpitch;
  float* rowstats;       // [B][heads][N][4]: max (log2 domain), 1 / sum, D, unused
  int C, N;
  float scale;
};

template <int DT>
__global__ void __launch_bounds__(kAbThreads) attn_bwd_dq_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(16) float s_kv[];   // K [N][8], V [N][8]
  float* s_k = s_kv;
  float* s_v = s_kv + (size_t)a.N * 8;
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = a.qkv + (size_t)b * a.N * a.pitch + head * 8;
  for (int j = threadIdx.x; j < a.N; j += kAbThreads) {
    float f[8];
    load8<DT>(base + (size_t)j * a.pitch + a.C, f);
#pragma unroll
    for (int d = 0; d < 8; ++d) s_k[j * 8 + d] = f[d];
    load8<DT>(base + (size_t)j * a.pitch + 2 * a.C, f);
#pragma unroll
    for (int d = 0; d < 8; ++d) s_v[j * 8 + d] = f[d];
  }
  __syncthreads();
  const int i = blockIdx.x * kAbThreads + threadIdx.x;
  if (i >= a.N) return;
  float q[8], go[8], ov[8];
  load8<DT>(base + (size_t)i * a.pitch, q);
  load8<DT>(a.d_o + ((size_t)b * a.N + i) * a.dopitch + head * 8, go);
  load8<DT>(a.o + ((size_t)b * a.N + i) * a.opitch + head * 8, ov);
  float D = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) D = fmaf(go[d], ov[d], D);
  const float sl = a.scale * kLog2e;
  float qs[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) qs[d] = q[d] * sl;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < a.N; ++j) {
    const float4 k0 = *reinterpret_cast<const float4*>(&s_k[j * 8]);
    const float4 k1 = *reinterpret_cast<const float4*>(&s_k[j * 8 + 4]);
    float s = qs[0] * k0.x;
    s = fmaf(qs[1], k0.y, s); s = fmaf(qs[2], k0.z, s); s = fmaf(qs[3], k0.w, s);
    s = fmaf(qs[4], k1.x, s); s = fmaf(qs[5], k1.y, s); s = fmaf(qs[6], k1.z, s); s = fmaf(qs[7], k1.w, s);
    if (s > m) {
      l *= ex2(m - s);
      m = s;
    }
    l += ex2(s - m);
  }
  const float inv_l = 1.0f / l;
  float dq[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) dq[d] = 0.f;
  for (int j = 0; j < a.N; ++j) {
    const float4 k0 = *reinterpret_cast<const float4*>(&s_k[j * 8]);
    const float4 k1 = *reinterpret_cast<const float4*>(&s_k[j * 8 + 4]);
    const float4 v0 = *reinterpret_cast<const float4*>(&s_v[j * 8]);
    const float4 v1 = *reinterpret_cast<const float4*>(&s_v[j * 8 + 4]);
    float s = qs[0] * k0.x;
    s = fmaf(qs[1], k0.y, s); s = fmaf(qs[2], k0.z, s); s = fmaf(qs[3], k0.w, s);
    s = fmaf(qs[4], k1.x, s); s = fmaf(qs[5], k1.y, s); s = fmaf(qs[6], k1.z, s); s = fmaf(qs[7], k1.w, s);
    const float p = ex2(s - m) * inv_l;
    float dp = go[0] * v0.x;
    dp = fmaf(go[1], v0.y, dp); dp = fmaf(go[2], v0.z, dp); dp = fmaf(go[3], v0.w, dp);
    dp = fmaf(go[4], v1.x, dp); dp = fmaf(go[5], v1.y, dp); dp = fmaf(go[6], v1.z, dp); dp = fmaf(go[7], v1.w, dp);
    const float ds = p * (dp - D);
    dq[0] = fmaf(ds, k0.x, dq[0]); dq[1] = fmaf(ds, k0.y, dq[1]); dq[2] = fmaf(ds, k0.z, dq[2]); dq[3] = fmaf(ds, k0.w, dq[3]);
    dq[4] = fmaf(ds, k1.x, dq[4]); dq[5] = fmaf(ds, k1.y, dq[5]); dq[6] = fmaf(ds, k1.z, dq[6]); dq[7] = fmaf(ds, k1.w, dq[7]);
  }
#pragma unroll
  for (int d = 0; d < 8; ++d) dq[d] *= a.scale;
  store8<DT>(a.dqkv + ((size_t)b * a.N + i) * a.dpitch + head * 8, dq);
  float4 st = make_float4(m, inv_l, D, 0.f);
  *reinterpret_cast<float4*>(a.rowstats + (((size_t)b * gridDim.y + head) * a.N + i) * 4) = st;
}

template <int DT>
__global__ void __launch_bounds__(kAbThreads) attn_bwd_dkv_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(16) float s_q[];   // per query: q*scale*log2e [8], d_o [8], stats [4] -> 20 floats
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = a.qkv + (size_t)b * a.N * a.pitch + head * 8;
  const float sl = a.scale * kLog2e;
  for (int i = threadIdx.x; i < a.N; i += kAbThreads) {
    float f[8];
    load8<DT>(base + (size_t)i * a.pitch, f);
#pragma unroll
    for (int d = 0; d < 8; ++d) s_q[i * 20 + d] = f[d] * sl;
    load8<DT>(a.d_o + ((size_t)b * a.N + i) * a.dopitch + head * 8, f);
#pragma unroll
    for (int d = 0; d < 8; ++d) s_q[i * 20 + 8 + d] = f[d];
    const float4 st = *reinterpret_cast<const float4*>(a.rowstats + (((size_t)b * gridDim.y + head) * a.N + i) * 4);
    *reinterpret_cast<float4*>(&s_q[i * 20 + 16]) = st;
  }
  __syncthreads();
  const int j = blockIdx.x * kAbThreads + threadIdx.x;
  if (j >= a.N) return;
  float k[8], v[8];
  load8<DT>(base + (size_t)j * a.pitch + a.C, k);
  load8<DT>(base + (size_t)j * a.pitch + 2 * a.C, v);
  float dk[8], dv[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) dk[d] = dv[d] = 0.f;
  for (int i = 0; i < a.N; ++i) {
    const float4 q0 = *reinterpret_cast<const float4*>(&s_q[i * 20]);
    const float4 q1 = *reinterpret_cast<const float4*>(&s_q[i * 20 + 4]);
    const float4 g0 = *reinterpret_cast<const float4*>(&s_q[i * 20 + 8]);
    const float4 g1 = *reinterpret_cast<const float4*>(&s_q[i * 20 + 12]);
    const float4 st = *reinterpret_cast<const float4*>(&s_q[i * 20 + 16]);
    float s = q0.x * k[0];
    s = fmaf(q0.y, k[1], s); s = fmaf(q0.z, k[2], s); s = fmaf(q0.w, k[3], s);
    s = fmaf(q1.x, k[4], s); s = fmaf(q1.y, k[5], s); s = fmaf(q1.z, k[6], s); s = fmaf(q1.w, k[7], s);
    const float p = ex2(s - st.x) * st.y;
    float dp = g0.x * v[0];
    dp = fmaf(g0.y, v[1], dp); dp = fmaf(g0.z, v[2], dp); dp = fmaf(g0.w, v[3], dp);
    dp = fmaf(g1.x, v[4], dp); dp = fmaf(g1.y, v[5], dp); dp = fmaf(g1.z, v[6], dp); dp = fmaf(g1.w, v[7], dp);
    const float ds = p * (dp - st.z);
    dv[0] = fmaf(p, g0.x, dv[0]); dv[1] = fmaf(p, g0.y, dv[1]); dv[2] = fmaf(p, g0.z, dv[2]); dv[3] = fmaf(p, g0.w, dv[3]);
    dv[4] = fmaf(p, g1.x, dv[4]); dv[5] = fmaf(p, g1.y, dv[5]); dv[6] = fmaf(p, g1.z, dv[6]); dv[7] = fmaf(p, g1.w, dv[7]);
    dk[0] = fmaf(ds, q0.x, dk[0]); dk[1] = fmaf(ds, q0.y, dk[1]); dk[2] = fmaf(ds, q0.z, dk[2]); dk[3] = fmaf(ds, q0.w, dk[3]);
    dk[4] = fmaf(ds, q1.x, dk[4]); dk[5] = fmaf(ds, q1.y, dk[5]); dk[6] = fmaf(ds, q1.z, dk[6]); dk[7] = fmaf(ds, q1.w, dk[7]);
  }
  // s_q holds q * scale * log2e: dk = scale * sum dS q = (1 / log2e) * sum dS (q * scale * log2e)
  const float un = 1.0f / kLog2e;
#pragma unroll
  for (int d = 0; d < 8; ++d) dk[d] *= un;
  store8<DT>(a.dqkv + ((size_t)b * a.N + j) * a.dpitch + a.C + head * 8, dk);
  store8<DT>(a.dqkv + ((size_t)b * a.N + j) * a.dpitch + 2 * a.C + head * 8, dv);
}

}  // namespace
}  // namespace d3

extern "C" size_t d3_attention_hd8_bwd_workspace_bytes(int B, int N, int C) {
  return (size_t)B * (C / 8) * N * 4 * sizeof(float);
}

extern "C" int d3_attention_hd8_bwd(const void* qkv, int pitch, const void* o, int opitch, const void* d_o, int dopitch,
                                    int B, int N, int C, void* dqkv, int dpitch, float* workspace, int dtype,
                                    void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(qkv && o && d_o && dqkv && workspace, "attention_bwd: null pointer");
  D3_REQUIRE(C % 8 == 0 && pitch % 8 == 0 && opitch % 8 == 0 && dopitch % 8 == 0 && dpitch % 8 == 0 && N > 0,
             "attention_bwd: bad shapes");
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "attention_bwd: 16-bit dtypes only");
  const size_t smem_a = (size_t)N * 16 * sizeof(float), smem_b = (size_t)N * 20 * sizeof(float);
  D3_REQUIRE(smem_b <= 220 * 1024, "attention_bwd: N = %d tokens do not fit shared memory", N);
  AttnBwdArgs a{};
  a.qkv = static_cast<const uint16_t*>(qkv);
  a.pitch = pitch;
  a.o = static_cast<const uint16_t*>(o);
  a.opitch = opitch;
  a.d_o = static_cast<const uint16_t*>(d_o);
  a.dopitch = dopitch;
  a.dqkv = static_cast<uint16_t*>(dqkv);
  a.dpitch = dpitch;
  a.rowstats = workspace;
  a.C = C;
  a.N = N;
  a.scale = 0.35355339059327373f;   // 1 / sqrt(8)
  dim3 grid(ceil_div(N, kAbThreads), C / 8, B);
  if (dtype == D3_DTYPE_F16) {
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_kernel<kF16>), smem_a));
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkv_kernel<kF16>), smem_b));
    attn_bwd_dq_kernel<kF16><<<grid, kAbThreads, smem_a, stream>>>(a);
    D3_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<kF16><<<grid, kAbThreads, smem_b, stream>>>(a);
  } else {
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_kernel<kBF16>), smem_a));
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkv_kernel<kBF16>), smem_b));
    attn_bwd_dq_kernel<kBF16><<<grid, kAbThreads, smem_a, stream>>>(a);
    D3_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<kBF16><<<grid, kAbThreads, smem_b, stream>>>(a);
  }
  D3_LAUNCH_CHECK();
  return D3_OK;
}
