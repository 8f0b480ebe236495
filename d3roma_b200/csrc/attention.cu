// Spatial self-attention core of the UNet attention blocks: 64 heads x head_dim 8,
// N = H*W tokens (960 at 24x40, 240 at 12x20, 2040 at 34x60), softmax(q k^T / sqrt(8)) v.
// Replaces aten::scaled_dot_product_attention as invoked by diffusers' Attention processor (blocks built at
// /root/reference/train.py:788-810).  The GroupNorm, the fused QKV projection and the output projection + residual
// run on the tcgen05 igemm kernel; this kernel is the softmax(QK^T)V part.
//
// Tensor-core formulation sized for head_dim = 8 (register-fragment MMA, flash-attention style):
//   S = Q K^T   : mma.m16n8k8   (k = head_dim = 8 exactly, no padding)      16 queries x 8 keys per instruction
//   O += P V    : mma.m16n8k16  (k = 16 keys, n = head_dim = 8)             the S accumulator fragments ARE the P
//                                                                            A-fragments (no shuffle, no smem trip)
// One warp owns 16 queries; K ([N][8]) and V^T ([8][N+pad]) of the (sample, head) live in shared memory (16 B/token
// each).  Online softmax in the exp2 domain over 64-key blocks.  The op is exp-bound (one MUFU.EX2 per score), not
// MMA-bound: a tcgen05/TMEM variant would have to pad K 8->16 and could not go faster than the MUFU pipe either; see
// DESIGN.md.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

// Warps (x 16 queries) per block: every block stages the whole K / V of its (sample, head) in shared memory, so more
// warps per block means fewer redundant K/V loads (N = 960: 8 blocks of 128 queries instead of 15 of 64).
constexpr int kHeadDim = 8;
constexpr int kKeyBlock = 64;
constexpr int kVPad = 8;  // V^T row pitch = N + 8 elements -> conflict-free fragment loads

D3_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DT>
D3_DEVINL void mma_m16n8k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
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
D3_DEVINL void mma_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                            uint32_t b1) {
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

template <int DT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attention_hd8_kernel(const uint16_t* __restrict__ qkv, int pitch,
                                                                     int C, int N, uint16_t* __restrict__ out,
                                                                     int opitch) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int Npad = (N + kKeyBlock - 1) / kKeyBlock * kKeyBlock;   // keys padded to a whole block (masked)
  const int vpitch = Npad + kVPad;                                // elements per V^T row
  uint4* s_k = reinterpret_cast<uint4*>(s_raw);                   // [Npad] rows of 8 x 16-bit
  uint16_t* s_vt = reinterpret_cast<uint16_t*>(s_raw + (size_t)Npad * 16);  // [8][vpitch]
  const int head = blockIdx.y, b = blockIdx.z;
  const uint16_t* base = qkv + (size_t)b * N * pitch + head * kHeadDim;
  constexpr int kAttnThreads = WARPS * 32;
  constexpr int kQPerBlock = WARPS * 16;
  for (int j = threadIdx.x; j < Npad; j += kAttnThreads) {
    uint4 kr = make_uint4(0u, 0u, 0u, 0u), vr = make_uint4(0u, 0u, 0u, 0u);
    if (j < N) {
      kr = __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * pitch + C));
      vr = __ldg(reinterpret_cast<const uint4*>(base + (size_t)j * pitch + 2 * C));
    }
    s_k[j] = kr;
    const uint32_t w[4] = {vr.x, vr.y, vr.z, vr.w};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      s_vt[(2 * d) * vpitch + j] = (uint16_t)(w[d] & 0xFFFFu);
      s_vt[(2 * d + 1) * vpitch + j] = (uint16_t)(w[d] >> 16);
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * kQPerBlock + warp * 16;
  if (q0 >= N) return;
  // Q fragment (A of m16n8k8): rows g and g+8, dims 2t, 2t+1
  uint32_t qa0 = 0u, qa1 = 0u;
  if (q0 + g < N) qa0 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g) * pitch + 2 * t));
  if (q0 + g + 8 < N) qa1 = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g + 8) * pitch + 2 * t));

  const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t* s_k32 = reinterpret_cast<const uint32_t*>(s_k);

  // One 64-key block of the online softmax.  MASKED: keys >= N (only the last, partial block has them) are set to -inf;
  // the full blocks skip that test - it was 23 % of the kernel's instructions (ncu source page, profiles/r02_*).
  auto key_block = [&](const int kb, auto masked_tag) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    float s[kKeyBlock / 8][4];
#pragma unroll
    for (int i = 0; i < kKeyBlock / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // B fragment: K[key = kb + 8i + g][dims 2t, 2t+1]
      const uint32_t kf = s_k32[(kb + 8 * i + g) * 4 + t];
      mma_m16n8k8<DT>(s[i], qa0, qa1, kf);
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < kKeyBlock / 8; ++i) {
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
    const float c0 = fast_exp2((m0 - mx0) * sc), c1 = fast_exp2((m1 - mx1) * sc);
    m0 = mx0;
    m1 = mx1;
    l0 *= c0;
    l1 *= c1;
    o[0] *= c0; o[1] *= c0; o[2] *= c1; o[3] *= c1;
    const float nm0 = -m0 * sc, nm1 = -m1 * sc;
#pragma unroll
    for (int i = 0; i < kKeyBlock / 16; ++i) {
      // 16 keys: S fragments 2i (keys +0..7) and 2i+1 (keys +8..15) become the A fragment of the PV MMA
      float p[8];
      p[0] = fast_exp2(fmaf(s[2 * i][0], sc, nm0));
      p[1] = fast_exp2(fmaf(s[2 * i][1], sc, nm0));
      p[2] = fast_exp2(fmaf(s[2 * i][2], sc, nm1));
      p[3] = fast_exp2(fmaf(s[2 * i][3], sc, nm1));
      p[4] = fast_exp2(fmaf(s[2 * i + 1][0], sc, nm0));
      p[5] = fast_exp2(fmaf(s[2 * i + 1][1], sc, nm0));
      p[6] = fast_exp2(fmaf(s[2 * i + 1][2], sc, nm1));
      p[7] = fast_exp2(fmaf(s[2 * i + 1][3], sc, nm1));
      l0 += (p[0] + p[1]) + (p[4] + p[5]);
      l1 += (p[2] + p[3]) + (p[6] + p[7]);
      const uint32_t a0 = pack2<DT>(p[0], p[1]), a1 = pack2<DT>(p[2], p[3]);
      const uint32_t a2 = pack2<DT>(p[4], p[5]), a3 = pack2<DT>(p[6], p[7]);
      // B fragment: V^T[dim g][keys kb+16i + 2t, +1] and [.. + 8 + 2t, +1]
      const uint16_t* vrow = s_vt + g * vpitch + kb + 16 * i + 2 * t;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vrow);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vrow + 8);
      mma_m16n8k16<DT>(o, a0, a1, a2, a3, b0, b1);
    }
  };
  const int n_full = (N / kKeyBlock) * kKeyBlock;
  for (int kb = 0; kb < n_full; kb += kKeyBlock) key_block(kb, std::false_type{});
  if (n_full < Npad) key_block(n_full, std::true_type{});
  // row sums are spread over the 4 threads of a quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  if (q0 + g < N)
    *reinterpret_cast<uint32_t*>(out + ((size_t)b * N + q0 + g) * opitch + head * kHeadDim + 2 * t) =
        pack2<DT>(o[0] * i0, o[1] * i0);
  if (q0 + g + 8 < N)
    *reinterpret_cast<uint32_t*>(out + ((size_t)b * N + q0 + g + 8) * opitch + head * kHeadDim + 2 * t) =
        pack2<DT>(o[2] * i1, o[3] * i1);
}

}  // namespace
}  // namespace d3

extern "C" int d3_attention_hd8(const void* qkv, int pitch, int B, int N, int C, void* out, int opitch,
                                int dtype, void* stream) {
  using namespace d3;
  D3_REQUIRE(qkv && out && C % 8 == 0 && pitch >= 3 * C && pitch % 8 == 0 && opitch % 2 == 0,
             "attention: bad arguments");
  if (dtype == D3_DTYPE_F32) return attention_f32(qkv, pitch, B, N, C, out, opitch, static_cast<cudaStream_t>(stream));
  const int Npad = (N + kKeyBlock - 1) / kKeyBlock * kKeyBlock;
  const size_t smem = (size_t)Npad * 16 + (size_t)8 * (Npad + kVPad) * 2;
  D3_REQUIRE(smem <= 220 * 1024, "attention: N=%d tokens exceed the shared-memory K/V budget", N);
  static int warps = -1;
  if (warps < 0) { const char* e = getenv("D3_ATTN_WARPS"); warps = (e && atoi(e) == 4) ? 4 : 8; }
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint16_t* q = static_cast<const uint16_t*>(qkv);
  uint16_t* o = static_cast<uint16_t*>(out);
#define D3_ATTN_LAUNCH(DTV, WV)                                                                                 \
  do {                                                                                                           \
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attention_hd8_kernel<DTV, WV>), smem));          \
    dim3 grid(ceil_div(N, WV * 16), C / kHeadDim, B);                                                            \
    D3_CHECK_CUDA(launch_pdl(attention_hd8_kernel<DTV, WV>, grid, dim3(WV * 32), smem, st, q, pitch, C, N, o, opitch)); \
  } while (0)
  if (dtype == D3_DTYPE_F16) {
    if (warps == 8) D3_ATTN_LAUNCH(kF16, 8); else D3_ATTN_LAUNCH(kF16, 4);
  } else {
    if (warps == 8) D3_ATTN_LAUNCH(kBF16, 8); else D3_ATTN_LAUNCH(kBF16, 4);
  }
#undef D3_ATTN_LAUNCH
  D3_LAUNCH_CHECK();
  return D3_OK;
}
