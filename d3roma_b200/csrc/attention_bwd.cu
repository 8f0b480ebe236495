// Backward of the spatial self-attention core (64 heads x head_dim 8; forward: attention.cu):
// given q, k, v (the fused QKV projection output), the forward output o and the gradient d_o, produce dq, dk, dv in
// the layout of the QKV tensor.  Replaces the backward of aten::scaled_dot_product_attention in the reference's
// training step (/root/reference/train.py:455 `accelerator.backward(loss)`).
//
// Two flash-attention-style passes on register-fragment tensor-core MMAs (the formulation of the forward kernel: head_dim
// 8 is exactly the k of mma.m16n8k8 and the n of mma.m16n8k16), probabilities recomputed, no atomics:
//   attn_bwd_dq_kernel   a warp owns 16 QUERIES; K, V ([N][8]) and K^T ([8][N]) of the (sample, head) in shared memory.
//                        Per 8 keys: S = Q K^T and dP = dO V^T (m16n8k8 each); e = exp2(S - max) with the forward's
//                        online rescaling; dS~ = e (dP - D), D_i = dO_i . O_i; the S-shaped accumulator fragments of two
//                        key groups ARE the A fragment of dQ~ += dS~ K (m16n8k16).  dQ = scale dQ~ / rowsum.  The row
//                        statistics (max, 1 / sum, D) are parked in a small fp32 workspace.
//   attn_bwd_dkv_kernel  a warp owns 16 KEYS; Q, dO ([N][8]), Q^T, dO^T ([8][N]) and the row statistics of every query in
//                        shared memory.  S^T = K Q^T, dP^T = V dO^T (m16n8k8), P^T = exp2(S^T - max_q) / sum_q,
//                        dS^T = P^T (dP^T - D_q); dV += P^T dO, dK += dS^T Q (m16n8k16).
// Round-2 history: the first version did this on CUDA cores (one thread per query / key, ~40 FMAs per score): 18.2 ms of a
// 158 ms training step at batch 16; see profiles/r02_train_launches.txt.
#include <cmath>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kAbWarps = 8;
constexpr int kBlk = 64;       // keys (dq kernel) / queries (dkv kernel) per inner block
constexpr int kTPad = 8;       // transposed rows: pitch = Npad + 8 elements -> conflict-free fragment loads
constexpr float kLog2e = 1.4426950408889634f;

D3_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DT>
D3_DEVINL void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  if constexpr (DT == kBF16) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  } else {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
}
template <int DT>
D3_DEVINL void mma_k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
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

struct AttnBwdArgs {
  const uint16_t* qkv;   // [B, N, 3C]: q | k | v
  int pitch;
  const uint16_t* o;     // [B, N, C]
  int opitch;
  const uint16_t* d_o;   // [B, N, C]
  int dopitch;
  uint16_t* dqkv;        // [B, N, 3C]
  int dpitch;
  float* rowstats;       // [B][heads][N][4]: max * scale * log2e, 1 / sum, D, unused
  int C, N;
  float scale;
};

// stage rows [N][8] of a 16-bit tensor (row pitch `pitch`) as [Npad] uint4 rows and as the transposed [8][tp] matrix
D3_DEVINL void stage_rows(const uint16_t* src, int pitch, int N, int Npad, int tp, uint4* rows, uint16_t* trans, int nthreads) {
  for (int j = threadIdx.x; j < Npad; j += nthreads) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (j < N) r = __ldg(reinterpret_cast<const uint4*>(src + (size_t)j * pitch));
    if (rows != nullptr) rows[j] = r;
    if (trans != nullptr) {
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        trans[(2 * d) * tp + j] = (uint16_t)(w[d] & 0xFFFFu);
        trans[(2 * d + 1) * tp + j] = (uint16_t)(w[d] >> 16);
      }
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(kAbWarps * 32) attn_bwd_dq_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int N = a.N;
  const int Npad = (N + kBlk - 1) / kBlk * kBlk;
  const int tp = Npad + kTPad;
  uint4* s_k = reinterpret_cast<uint4*>(s_raw);
  uint4* s_v = s_k + Npad;
  uint16_t* s_kt = reinterpret_cast<uint16_t*>(s_v + Npad);
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = a.qkv + (size_t)b * N * a.pitch + head * 8;
  stage_rows(base + a.C, a.pitch, N, Npad, tp, s_k, s_kt, kAbWarps * 32);
  stage_rows(base + 2 * a.C, a.pitch, N, Npad, tp, s_v, nullptr, kAbWarps * 32);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * (kAbWarps * 16) + warp * 16;
  if (q0 >= N) return;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  uint32_t qa0 = 0u, qa1 = 0u, ga0 = 0u, ga1 = 0u, oa0 = 0u, oa1 = 0u;
  if (r0 < N) {
    qa0 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r0 * a.pitch + 2 * t));
    ga0 = __ldg(reinterpret_cast<const uint32_t*>(a.d_o + ((size_t)b * N + r0) * a.dopitch + head * 8 + 2 * t));
    oa0 = __ldg(reinterpret_cast<const uint32_t*>(a.o + ((size_t)b * N + r0) * a.opitch + head * 8 + 2 * t));
  }
  if (r1 < N) {
    qa1 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r1 * a.pitch + 2 * t));
    ga1 = __ldg(reinterpret_cast<const uint32_t*>(a.d_o + ((size_t)b * N + r1) * a.dopitch + head * 8 + 2 * t));
    oa1 = __ldg(reinterpret_cast<const uint32_t*>(a.o + ((size_t)b * N + r1) * a.opitch + head * 8 + 2 * t));
  }
  float D0, D1;
  {
    const float2 gg0 = unpack2<DT>(ga0), oo0 = unpack2<DT>(oa0), gg1 = unpack2<DT>(ga1), oo1 = unpack2<DT>(oa1);
    D0 = gg0.x * oo0.x + gg0.y * oo0.y;
    D1 = gg1.x * oo1.x + gg1.y * oo1.y;
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1);
    D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  }
  const float sc = a.scale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float dq[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t* s_k32 = reinterpret_cast<const uint32_t*>(s_k);
  const uint32_t* s_v32 = reinterpret_cast<const uint32_t*>(s_v);

  auto key_block = [&](const int kb, auto masked_tag) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    float s[kBlk / 8][4], dp[kBlk / 8][4];
#pragma unroll
    for (int i = 0; i < kBlk / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      mma_k8<DT>(s[i], qa0, qa1, s_k32[(kb + 8 * i + g) * 4 + t]);
      mma_k8<DT>(dp[i], ga0, ga1, s_v32[(kb + 8 * i + g) * 4 + t]);
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < kBlk / 8; ++i) {
      if constexpr (MASKED) {
        const int key = kb + 8 * i + 2 * t;
        if (key >= N) s[i][0] = s[i][2] = -INFINITY;
        if (key + 1 >= N) s[i][1] = s[i][3] = -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ex2((m0 - mx0) * sc), c1 = ex2((m1 - mx1) * sc);
    m0 = mx0;
    m1 = mx1;
    l0 *= c0;
    l1 *= c1;
    dq[0] *= c0; dq[1] *= c0; dq[2] *= c1; dq[3] *= c1;
    const float nm0 = -m0 * sc, nm1 = -m1 * sc;
#pragma unroll
    for (int i = 0; i < kBlk / 16; ++i) {
      float e[8];
      e[0] = ex2(fmaf(s[2 * i][0], sc, nm0));
      e[1] = ex2(fmaf(s[2 * i][1], sc, nm0));
      e[2] = ex2(fmaf(s[2 * i][2], sc, nm1));
      e[3] = ex2(fmaf(s[2 * i][3], sc, nm1));
      e[4] = ex2(fmaf(s[2 * i + 1][0], sc, nm0));
      e[5] = ex2(fmaf(s[2 * i + 1][1], sc, nm0));
      e[6] = ex2(fmaf(s[2 * i + 1][2], sc, nm1));
      e[7] = ex2(fmaf(s[2 * i + 1][3], sc, nm1));
      l0 += (e[0] + e[1]) + (e[4] + e[5]);
      l1 += (e[2] + e[3]) + (e[6] + e[7]);
      // dS~ = e (dP - D): the two S-shaped fragments become the A fragment (16 queries x 16 keys) of dQ~ += dS~ K
      const uint32_t a0 = pack2<DT>(e[0] * (dp[2 * i][0] - D0), e[1] * (dp[2 * i][1] - D0));
      const uint32_t a1 = pack2<DT>(e[2] * (dp[2 * i][2] - D1), e[3] * (dp[2 * i][3] - D1));
      const uint32_t a2 = pack2<DT>(e[4] * (dp[2 * i + 1][0] - D0), e[5] * (dp[2 * i + 1][1] - D0));
      const uint32_t a3 = pack2<DT>(e[6] * (dp[2 * i + 1][2] - D1), e[7] * (dp[2 * i + 1][3] - D1));
      // B fragment: K^T[dim g][keys kb+16i + 2t, +1] and [.. + 8 + 2t, +1]
      const uint16_t* krow = s_kt + g * tp + kb + 16 * i + 2 * t;
      mma_k16<DT>(dq, a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(krow), *reinterpret_cast<const uint32_t*>(krow + 8));
    }
  };
  const int n_full = (N / kBlk) * kBlk;
  for (int kb = 0; kb < n_full; kb += kBlk) key_block(kb, std::false_type{});
  if (n_full < Npad) key_block(n_full, std::true_type{});
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  float* st = a.rowstats + ((size_t)b * gridDim.y + head) * N * 4;
  if (r0 < N) {
    *reinterpret_cast<uint32_t*>(a.dqkv + ((size_t)b * N + r0) * a.dpitch + head * 8 + 2 * t) =
        pack2<DT>(dq[0] * i0 * a.scale, dq[1] * i0 * a.scale);
    if (t == 0) *reinterpret_cast<float4*>(st + (size_t)r0 * 4) = make_float4(m0 * sc, i0, D0, 0.f);
  }
  if (r1 < N) {
    *reinterpret_cast<uint32_t*>(a.dqkv + ((size_t)b * N + r1) * a.dpitch + head * 8 + 2 * t) =
        pack2<DT>(dq[2] * i1 * a.scale, dq[3] * i1 * a.scale);
    if (t == 0) *reinterpret_cast<float4*>(st + (size_t)r1 * 4) = make_float4(m1 * sc, i1, D1, 0.f);
  }
}

template <int DT>
__global__ void __launch_bounds__(kAbWarps * 32) attn_bwd_dkv_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int N = a.N;
  const int Npad = (N + kBlk - 1) / kBlk * kBlk;
  const int tp = Npad + kTPad;
  uint4* s_q = reinterpret_cast<uint4*>(s_raw);
  uint4* s_g = s_q + Npad;
  uint16_t* s_qt = reinterpret_cast<uint16_t*>(s_g + Npad);
  uint16_t* s_gt = s_qt + 8 * tp;
  float* s_msc = reinterpret_cast<float*>(s_gt + 8 * tp);
  float* s_il = s_msc + Npad;
  float* s_D = s_il + Npad;
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = a.qkv + (size_t)b * N * a.pitch + head * 8;
  stage_rows(base, a.pitch, N, Npad, tp, s_q, s_qt, kAbWarps * 32);
  stage_rows(a.d_o + (size_t)b * N * a.dopitch + head * 8, a.dopitch, N, Npad, tp, s_g, s_gt, kAbWarps * 32);
  const float* st = a.rowstats + ((size_t)b * gridDim.y + head) * N * 4;
  for (int i = threadIdx.x; i < Npad; i += kAbWarps * 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // padded queries: 1 / sum = 0 -> they contribute nothing
    if (i < N) v = __ldg(reinterpret_cast<const float4*>(st + (size_t)i * 4));
    s_msc[i] = v.x;
    s_il[i] = v.y;
    s_D[i] = v.z;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * (kAbWarps * 16) + warp * 16;
  if (k0 >= N) return;
  const int r0 = k0 + g, r1 = k0 + g + 8;
  uint32_t ka0 = 0u, ka1 = 0u, va0 = 0u, va1 = 0u;
  if (r0 < N) {
    ka0 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r0 * a.pitch + a.C + 2 * t));
    va0 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r0 * a.pitch + 2 * a.C + 2 * t));
  }
  if (r1 < N) {
    ka1 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r1 * a.pitch + a.C + 2 * t));
    va1 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r1 * a.pitch + 2 * a.C + 2 * t));
  }
  const float sc = a.scale * kLog2e;
  float dk[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t* s_q32 = reinterpret_cast<const uint32_t*>(s_q);
  const uint32_t* s_g32 = reinterpret_cast<const uint32_t*>(s_g);
  for (int qb = 0; qb < Npad; qb += kBlk) {
    float s[kBlk / 8][4], dp[kBlk / 8][4];
#pragma unroll
    for (int i = 0; i < kBlk / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      // S^T = K Q^T: rows = this warp's keys, columns = queries qb + 8i + (0..7)
      mma_k8<DT>(s[i], ka0, ka1, s_q32[(qb + 8 * i + g) * 4 + t]);
      mma_k8<DT>(dp[i], va0, va1, s_g32[(qb + 8 * i + g) * 4 + t]);
    }
#pragma unroll
    for (int i = 0; i < kBlk / 16; ++i) {
      float p[8], ds[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = 2 * i + h;
        const int qq = qb + 8 * n + 2 * t;          // this lane's two query columns
        const float2 msc = *reinterpret_cast<const float2*>(&s_msc[qq]);
        const float2 il = *reinterpret_cast<const float2*>(&s_il[qq]);
        const float2 dd = *reinterpret_cast<const float2*>(&s_D[qq]);
        p[4 * h + 0] = ex2(fmaf(s[n][0], sc, -msc.x)) * il.x;
        p[4 * h + 1] = ex2(fmaf(s[n][1], sc, -msc.y)) * il.y;
        p[4 * h + 2] = ex2(fmaf(s[n][2], sc, -msc.x)) * il.x;
        p[4 * h + 3] = ex2(fmaf(s[n][3], sc, -msc.y)) * il.y;
        ds[4 * h + 0] = p[4 * h + 0] * (dp[n][0] - dd.x);
        ds[4 * h + 1] = p[4 * h + 1] * (dp[n][1] - dd.y);
        ds[4 * h + 2] = p[4 * h + 2] * (dp[n][2] - dd.x);
        ds[4 * h + 3] = p[4 * h + 3] * (dp[n][3] - dd.y);
      }
      // B fragments: dO^T / Q^T [dim g][queries qb+16i + 2t, +1] and [.. + 8 + 2t, +1]
      const uint16_t* grow = s_gt + g * tp + qb + 16 * i + 2 * t;
      const uint16_t* qrow = s_qt + g * tp + qb + 16 * i + 2 * t;
      mma_k16<DT>(dv, pack2<DT>(p[0], p[1]), pack2<DT>(p[2], p[3]), pack2<DT>(p[4], p[5]), pack2<DT>(p[6], p[7]),
                  *reinterpret_cast<const uint32_t*>(grow), *reinterpret_cast<const uint32_t*>(grow + 8));
      mma_k16<DT>(dk, pack2<DT>(ds[0], ds[1]), pack2<DT>(ds[2], ds[3]), pack2<DT>(ds[4], ds[5]), pack2<DT>(ds[6], ds[7]),
                  *reinterpret_cast<const uint32_t*>(qrow), *reinterpret_cast<const uint32_t*>(qrow + 8));
    }
  }
  if (r0 < N) {
    uint16_t* d = a.dqkv + ((size_t)b * N + r0) * a.dpitch + head * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(d + a.C) = pack2<DT>(dk[0] * a.scale, dk[1] * a.scale);
    *reinterpret_cast<uint32_t*>(d + 2 * a.C) = pack2<DT>(dv[0], dv[1]);
  }
  if (r1 < N) {
    uint16_t* d = a.dqkv + ((size_t)b * N + r1) * a.dpitch + head * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(d + a.C) = pack2<DT>(dk[2] * a.scale, dk[3] * a.scale);
    *reinterpret_cast<uint32_t*>(d + 2 * a.C) = pack2<DT>(dv[2], dv[3]);
  }
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
  const int Npad = (N + kBlk - 1) / kBlk * kBlk;
  const size_t tbytes = (size_t)8 * (Npad + kTPad) * 2;
  const size_t smem_a = (size_t)Npad * 32 + tbytes, smem_b = (size_t)Npad * 32 + 2 * tbytes + (size_t)Npad * 12;
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
  dim3 grid(ceil_div(N, kAbWarps * 16), C / 8, B);
  constexpr int kAbThreads = kAbWarps * 32;
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
