// Self-attention with head_dim 64 on the Blackwell tensor path: S = Q K^T and O = P V are tcgen05.mma instructions with
// TMEM accumulators, Q / K / V^T tiles arrive by TMA (SWIZZLE_128B, K-major), the softmax warps read S straight out of
// TMEM (one query row per thread: row max / row sum need no shuffles), write P as the 16-bit A operand of the second MMA
// into swizzled shared memory, and fold each tile's P V into register-resident output rows.
//
// Replaces aten::scaled_dot_product_attention of the transformer blocks' attn1 in diffusers' UNet2DConditionModel
// (SD-2.1 layout: 5 / 10 / 20 heads of 64 channels over 3600 / 920 / 240 / 60 tokens; latent family, loaded at
// /root/reference/train.py:743-772, called at core/custom_pipelines.py:386).
//
// One CTA = 128 queries of one (sample, head); 320 threads:
//   warp 0      TMEM allocation, then one lane = TMA producer (Q once; per 128-key tile K [128 x 64], double-buffered,
//               and V^T as two [64 dims x 64 keys] K-blocks)
//   warp 1      one lane = MMA issuer: QK (M 128, N 128, K 64) -> S in TMEM columns [0, 128);
//               PV (M 128, N 64, K 128) -> D in TMEM columns [128, 192)
//   warps 2-9   softmax: TWO threads per query row (TMEM lane 32 (warp % 4) + lane; warps 2-5 take key columns 0-63 of a
//               tile, warps 6-9 columns 64-127; the pair agrees on the row maximum through shared memory).  S is read
//               from TMEM into registers once, so the NEXT tile's QK runs under this tile's exponentials; O accumulates
//               in TMEM and is rescaled only when a row maximum outgrows its stabiliser by 2^8.
// 99 KB of shared memory and 256 TMEM columns per CTA: two CTAs (16 softmax warps) per SM.
// The op is bound by the exponentials (one MUFU.EX2 per score, 16 per clock and SM) and their latency, not by the tensor
// pipe (ncu: profiles/r02_ncu_attention_tc.txt).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "tensormap.h"

namespace d3 {
namespace {

constexpr int kTcThreads = 192;
constexpr uint32_t kQBytes = 128 * 128;       // 128 queries x 64 dims
constexpr uint32_t kKBytes = 128 * 128;       // 128 keys x 64 dims
constexpr uint32_t kVBlkBytes = 64 * 128;     // 64 dims x 64 keys
constexpr uint32_t kTcSmem = kQBytes + 2 * kKBytes + 2 * 2 * kVBlkBytes;   // 80 KB: Q + two stages of K and V^T
constexpr uint32_t kTcTmemCols = 256;

#ifdef D3_DEV
#define D3_ADBG(a, mask) ((a).dbg & (mask))
#else
#define D3_ADBG(a, mask) 0
#endif

#ifdef D3_DEV
// developer profiling (D3_ATTN_DBG bit 256): clock cycles summed over CTAs
//  softmax warp 2 lane 0: [0] total loop, [1] wait s_full, [2] S load, [3] max + exchange, [4] wait d_full, [5] rescale,
//                         [6] exps + pack, [7] P store + arrive;  MMA thread: [8] total, [9] wait s_free, [10] wait k_full,
//                         [11] wait p_full, [12] wait v_full;  [15] tiles
__device__ unsigned long long g_attn_prof[16];
#define D3_APROF(stmt) do { if (prof) { stmt; } } while (0)
#else
#define D3_APROF(stmt) do { } while (0)
#endif

struct AttnTcArgs {
  int dbg;   // D3_DEV builds only (D3_ATTN_DBG): skip phases to time the rest; results are garbage
  int N, ntiles;
  float scale_log2;
  uint16_t* out;
  int op;
};

D3_DEVINL float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DT>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_hd64_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + 2 * kKBytes;                 // [stage][key block of 64][64 dims x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 4 * kVBlkBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]  K and V^T tiles are double-buffered: the load of tile j+1 is requested a whole tile
  uint64_t* k_empty = bars + 3;  // [2]  ahead (with one buffer it can only start when tile j's MMA has completed, and the
  uint64_t* v_full = bars + 5;   // [2]  TMA latency - microseconds under load - sits on the critical path of every tile)
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // QK of a tile has completed: S is in TMEM
  uint64_t* s_free = bars + 10;  // the eight softmax warps hold S in registers: the next QK may overwrite it
  uint64_t* p_full = bars + 11;  // P of a tile is in TMEM (and D has been rescaled if it had to be)
  uint64_t* d_full = bars + 12;  // PV of a tile has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 128;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 13; ++i) mbar_init(&bars[i], (i == 10 || i == 11) ? 4 : 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, kTcTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_prologue();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, head * 64, q0, b);
      auto load_k = [&](int j) {
        const int buf = j & 1, use = j >> 1;      // use-th fill of this buffer
        if (use > 0) mbar_wait(&k_empty[buf], (uint32_t)(use - 1) & 1u);
        if (D3_ADBG(a, 64)) { mbar_arrive(&k_full[buf]); return; }
        mbar_arrive_expect_tx(&k_full[buf], kKBytes);
        tma_load_3d(sK + buf * kKBytes, &tmK, &k_full[buf], head * 64, j * 128, b);
      };
      auto load_v = [&](int j) {
        const int buf = j & 1, use = j >> 1;
        if (use > 0) mbar_wait(&v_empty[buf], (uint32_t)(use - 1) & 1u);
        if (D3_ADBG(a, 64)) { mbar_arrive(&v_full[buf]); return; }
        mbar_arrive_expect_tx(&v_full[buf], 2 * kVBlkBytes);
        uint8_t* dst = sV + buf * 2 * kVBlkBytes;
        tma_load_3d(dst, &tmV, &v_full[buf], j * 128, head * 64, b);
        tma_load_3d(dst + kVBlkBytes, &tmV, &v_full[buf], j * 128 + 64, head * 64, b);
      };
      load_k(0);
      load_v(0);
      for (int j = 1; j < a.ntiles; ++j) {
        load_k(j);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // Tensor-pipe order: QK(0), QK(1), PV(0), QK(2), PV(1), ... - QK(j+1) is issued as soon as the softmax warps have S(j)
    // in registers, so it runs under their exponentials; PV(j) follows when P(j) is in shared memory.
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_16bit(128, 128, DT == kBF16);
      constexpr uint32_t idesc_pv = make_idesc_16bit(128, 64, DT == kBF16);
      const uint64_t dQ = make_sw128_kmajor_desc(smem_u32(sQ));
      const uint64_t dK[2] = {make_sw128_kmajor_desc(smem_u32(sK)), make_sw128_kmajor_desc(smem_u32(sK + kKBytes))};
      uint64_t dV[2][2];
#pragma unroll
      for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) dV[st][kb] = make_sw128_kmajor_desc(smem_u32(sV + (st * 2 + kb) * kVBlkBytes));
      auto issue_qk = [&](int j) {
        const int buf = j & 1;
        mbar_wait(&k_full[buf], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!D3_ADBG(a, 16)) umma_bf16_ss(tmem_base, dQ + 2 * k, dK[buf] + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[buf]);
        umma_commit(s_full);
      };
#ifdef D3_DEV
      const bool prof = D3_ADBG(a, 256) != 0;
      long long m_t0 = prof ? clock64() : 0, m_sf = 0, m_kf = 0, m_pf = 0, m_vf = 0, m_c = 0;
#endif
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < a.ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        if (j + 1 < a.ntiles) {
          D3_APROF(m_c = clock64());
          mbar_wait(s_free, ph);
          D3_APROF(m_sf += clock64() - m_c; m_c = clock64());
          mbar_wait(&k_full[(j + 1) & 1], (uint32_t)((j + 1) >> 1) & 1u);
          D3_APROF(m_kf += clock64() - m_c);
          issue_qk(j + 1);
        }
        D3_APROF(m_c = clock64());
        mbar_wait(p_full, ph);
        D3_APROF(m_pf += clock64() - m_c; m_c = clock64());
        mbar_wait(&v_full[j & 1], (uint32_t)(j >> 1) & 1u);
        D3_APROF(m_vf += clock64() - m_c);
        tc_fence_after();
        // O (+)= P V: P is the A operand in TMEM columns [128, 192) (two keys per column), 8 K-steps of 16 keys
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!D3_ADBG(a, 8))
            umma_bf16_ts(tmem_base + 192, tmem_base + 128 + 8 * k, dV[j & 1][k >> 2] + 2 * (k & 3), idesc_pv,
                         (k != 0 || j > 0) ? 1u : 0u);
        umma_commit(&v_empty[j & 1]);
        umma_commit(d_full);
      }
#ifdef D3_DEV
      if (prof) {
        atomicAdd(&g_attn_prof[8], (unsigned long long)(clock64() - m_t0));
        atomicAdd(&g_attn_prof[9], (unsigned long long)m_sf);
        atomicAdd(&g_attn_prof[10], (unsigned long long)m_kf);
        atomicAdd(&g_attn_prof[11], (unsigned long long)m_pf);
        atomicAdd(&g_attn_prof[12], (unsigned long long)m_vf);
      }
#endif
    }
  } else {
    // ================================ softmax / output ================================
    const uint32_t q = warp & 3u;                    // TMEM lane quarter this warp may access
    const uint32_t row = q * 32 + lane;              // query row of this thread inside the tile
    const uint32_t trow = tmem_base + ((q * 32) << 16);
    const float sl2 = a.scale_log2;
    float m_used = -INFINITY, l = 0.f;
#ifdef D3_DEV
    const bool prof = D3_ADBG(a, 256) != 0 && threadIdx.x == 64;
    long long s_t0 = prof ? clock64() : 0, s_p[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s_c = 0;
#define D3_SPROF(i) D3_APROF(const long long n_ = clock64(); s_p[i] += n_ - s_c; s_c = n_)
#else
#define D3_SPROF(i) do { } while (0)
#endif
    for (int j = 0; j < a.ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      const int nvalid = a.N - j * 128;              // valid key columns of this tile (>= 128: all)
      D3_APROF(s_c = clock64());
      mbar_wait(s_full, ph);
      D3_SPROF(1);
      tc_fence_after();
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(trow + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);            // S(j) is in registers: QK(j+1) may run now
      D3_SPROF(2);
      float mx;
      if (D3_ADBG(a, 4)) {
        mx = 1.0f;
      } else if (nvalid >= 128) {
        float mq[4];                                 // four independent chains of 3-input maxima
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mq[c] = fmaxf(__uint_as_float(v[c][0]), __uint_as_float(v[c][1]));
#pragma unroll
          for (int i = 2; i < 32; i += 2) mq[c] = fmaxf(fmaxf(mq[c], __uint_as_float(v[c][i])), __uint_as_float(v[c][i + 1]));
        }
        mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
      } else {
        mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[c][i]));
      }
      D3_SPROF(3);
      // PV(j-1) must be complete before P is overwritten (and before D is rescaled)
      if (j > 0) {
        mbar_wait(d_full, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      D3_SPROF(4);
      const bool grow = (mx - m_used) * sl2 > 8.0f;                 // true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_used, mx);
        const float corr = ex2f((m_used - m_new) * sl2);            // 0 on the first tile
        m_used = m_new;
        l *= corr;
        if (j > 0) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t d[32];
            tmem_ld_32x32b_x32(trow + 192 + c * 32, d);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = __float_as_uint(__uint_as_float(d[i]) * corr);
            tmem_st_32x32b_x32(trow + 192 + c * 32, d);
          }
          tmem_st_wait();
        }
      }
      D3_SPROF(5);
      // P = exp2(S * scale - m_used * scale), two per 32-bit TMEM column (the A operand of the PV MMA); packed f32x2
      // arithmetic: one FFMA2 + two MUFU.EX2 + one FADD2 + one F2FP per pair of scores
      const float2 sl2v = make_float2(sl2, sl2), nmv = make_float2(-m_used * sl2, -m_used * sl2);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      auto exp_pass = [&](auto masked_tag) {       // MASKED: only the last, partial key tile pays for the column tests
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = make_float2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1]));
          asm("{\n\t.reg .b64 a, b, c;\n\t"
              "mov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tmov.b64 c, {%4, %5};\n\t"
              "fma.rn.f32x2 a, a, b, c;\n\t"
              "mov.b64 {%0, %1}, a;\n\t}"
              : "+f"(x.x), "+f"(x.y)
              : "f"(sl2v.x), "f"(sl2v.y), "f"(nmv.x), "f"(nmv.y));
          if (!D3_ADBG(a, 1)) {
            x.x = ex2f(x.x);
            x.y = ex2f(x.y);
          }
          if constexpr (MASKED) {
            if (c * 32 + 2 * i >= nvalid) x.x = 0.f;
            if (c * 32 + 2 * i + 1 >= nvalid) x.y = 0.f;
          }
          float2& ac = acc[i & 3];
          asm("{\n\t.reg .b64 a, b;\n\t"
              "mov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\t"
              "add.rn.f32x2 a, a, b;\n\t"
              "mov.b64 {%0, %1}, a;\n\t}"
              : "+f"(ac.x), "+f"(ac.y)
              : "f"(x.x), "f"(x.y));
          pk[i] = pack2<DT>(x.x, x.y);
        }
        // 16 columns = 32 keys of this row
        if (!D3_ADBG(a, 2))
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
              ::"r"(trow + 128 + c * 16), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]),
              "r"(pk[7]), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
              : "memory");
      }
      };
      if (nvalid >= 128) exp_pass(std::false_type{}); else exp_pass(std::true_type{});
      l += ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + ((acc[2].x + acc[2].y) + (acc[3].x + acc[3].y));
      D3_SPROF(6);
      tmem_st_wait();
      // P written (and D rescaled): hand over to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      D3_SPROF(7);
    }
#ifdef D3_DEV
    if (prof) {
      atomicAdd(&g_attn_prof[0], (unsigned long long)(clock64() - s_t0));
      for (int i = 1; i < 8; ++i) atomicAdd(&g_attn_prof[i], (unsigned long long)s_p[i]);
      atomicAdd(&g_attn_prof[15], (unsigned long long)a.ntiles);
    }
#endif
    mbar_wait(d_full, (uint32_t)(a.ntiles - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint16_t* dst = a.out + ((size_t)b * a.N + q0 + row) * a.op + head * 64;
    const bool valid = q0 + (int)row < a.N;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t d[32];
      tmem_ld_32x32b_x32(trow + 192 + c * 32, d);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack2<DT>(__uint_as_float(d[8 * i + 0]) * inv, __uint_as_float(d[8 * i + 1]) * inv);
          w.y = pack2<DT>(__uint_as_float(d[8 * i + 2]) * inv, __uint_as_float(d[8 * i + 3]) * inv);
          w.z = pack2<DT>(__uint_as_float(d[8 * i + 4]) * inv, __uint_as_float(d[8 * i + 5]) * inv);
          w.w = pack2<DT>(__uint_as_float(d[8 * i + 6]) * inv, __uint_as_float(d[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + 8 * i) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// vt[b][c][n] = v[b][n][c] for n < N, 0 for N <= n < Np: V^T with the keys contiguous (the K-major B operand of O = P V)
__global__ void __launch_bounds__(256) transpose_v_kernel(const uint16_t* __restrict__ v, int N, int C, int vp,
                                                          uint16_t* __restrict__ vt, int Np) {
  pdl_prologue();
  __shared__ uint16_t tile[32][34];
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint16_t* src = v + (size_t)b * N * vp;
  uint16_t* dst = vt + (size_t)b * C * Np;
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + tx;
    tile[i][tx] = (n < N && c < C) ? src[(size_t)n * vp + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    if (c < C && n < Np) dst[(size_t)c * Np + n] = tile[tx][i];
  }
}

}  // namespace
}  // namespace d3

using namespace d3;

extern "C" size_t d3_attention_hd64_tc_workspace_bytes(int B, int N, int heads) {
  const size_t Np = ((size_t)N + 127) / 128 * 128;   // whole key tiles: every TMA box starts inside the tensor
  return (size_t)B * heads * 64 * Np * 2;
}

extern "C" int d3_attention_hd64_tc(const void* q, int q_pitch, const void* k, int k_pitch, const void* v, int v_pitch,
                                    int B, int N, int heads, void* out, int out_pitch, void* workspace, int dtype,
                                    void* stream) {
  D3_REQUIRE(dtype == D3_DTYPE_BF16 || dtype == D3_DTYPE_F16, "attention_hd64_tc: 16-bit dtypes only (got %d)", dtype);
  D3_REQUIRE(q && k && v && out && workspace && B > 0 && N > 0 && heads > 0, "attention_hd64_tc: bad arguments");
  D3_REQUIRE(q_pitch % 8 == 0 && k_pitch % 8 == 0 && v_pitch % 8 == 0 && out_pitch % 8 == 0 &&
                 ((uintptr_t)q | (uintptr_t)k | (uintptr_t)out | (uintptr_t)workspace) % 16 == 0,
             "attention_hd64_tc: pitches must be multiples of 8 elements and pointers 16-byte aligned");
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = heads * 64;
  const int Np = (N + 127) / 128 * 128;
  // 1. V -> V^T (keys contiguous)
  {
    dim3 grid(ceil_div(Np, 32), ceil_div(C, 32), B);
    D3_CHECK_CUDA(launch_pdl(transpose_v_kernel, grid, dim3(256), 0, st, static_cast<const uint16_t*>(v), N, C, v_pitch,
                             static_cast<uint16_t*>(workspace), Np));
    D3_LAUNCH_CHECK();
  }
  // 2. tensor maps: Q / K rows of [B][N][pitch] (channel box = one head), V^T [B][C][Np]
  const bool f16 = dtype == D3_DTYPE_F16;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)B};
    uint32_t box[3] = {64, 128, 1};
    uint64_t sq[2] = {(uint64_t)q_pitch * 2, (uint64_t)N * q_pitch * 2};
    uint64_t sk[2] = {(uint64_t)k_pitch * 2, (uint64_t)N * k_pitch * 2};
    int rc;
    if ((rc = make_tmap_16bit(f16, &tmQ, q, 3, dims, sq, box))) return set_error(D3_ERR_CUDA, "attention_hd64_tc: tensor map Q (%d)", rc);
    if ((rc = make_tmap_16bit(f16, &tmK, k, 3, dims, sk, box))) return set_error(D3_ERR_CUDA, "attention_hd64_tc: tensor map K (%d)", rc);
    uint64_t dv[3] = {(uint64_t)Np, (uint64_t)C, (uint64_t)B};
    uint64_t sv[2] = {(uint64_t)Np * 2, (uint64_t)C * Np * 2};
    uint32_t bv[3] = {64, 64, 1};
    if ((rc = make_tmap_16bit(f16, &tmV, workspace, 3, dv, sv, bv))) return set_error(D3_ERR_CUDA, "attention_hd64_tc: tensor map V (%d)", rc);
  }
  int dbg = 0;
#ifdef D3_DEV
  { const char* e = getenv("D3_ATTN_DBG"); dbg = e ? atoi(e) : 0; }
#endif
  AttnTcArgs a{dbg, N, ceil_div(N, 128), 0.125f * 1.4426950408889634f, static_cast<uint16_t*>(out), out_pitch};
  const size_t smem = 1024 + kTcSmem + 128;   // alignment slack + tiles + barriers + row-max exchange
  dim3 grid(ceil_div(N, 128), heads, B);
  if (f16) {
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attention_hd64_tc_kernel<kF16>), smem));
    D3_CHECK_CUDA(launch_pdl(attention_hd64_tc_kernel<kF16>, grid, dim3(kTcThreads), smem, st, tmQ, tmK, tmV, a));
  } else {
    D3_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(attention_hd64_tc_kernel<kBF16>), smem));
    D3_CHECK_CUDA(launch_pdl(attention_hd64_tc_kernel<kBF16>, grid, dim3(kTcThreads), smem, st, tmQ, tmK, tmV, a));
  }
  D3_LAUNCH_CHECK();
  return D3_OK;
}

#ifdef D3_DEV
// Developer hook (D3_DEV builds only; not part of include/d3roma_b200.h)
extern "C" int d3_debug_attn_profile(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, d3::g_attn_prof, sizeof(d3::g_attn_prof)) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(d3::g_attn_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif
