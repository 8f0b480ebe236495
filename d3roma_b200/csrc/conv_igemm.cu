// Implicit-GEMM convolution for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
// GEMM view: M = output pixels (B*Ho*Wo), N = Cout, K = taps*Cin (+ folded 1x1
// shortcut channels). Activations are bf16 NHWC, so one K-block (64 channels of one filter tap) of the
// A operand for a TH x TW pixel tile is a single 4-D TMA box [64c, TW, TH, 1] whose start coordinate is
// shifted by the tap offset; out-of-image coordinates are zero-filled by TMA, which *is* the conv zero
// padding. Weights are pre-packed [Cout][Ktot] (K-major) so the B operand is a 2-D box [64k, BN].
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer (one lane)          smem ring: STAGES x {A 16 KB, B BN*128 B}, SWIZZLE_128B
//   warp 1   tcgen05.mma issuer (one lane)    accumulators: 2 x [128 lanes x BN cols] fp32 in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> +bias(+temb) (+identity residual) -> bf16 -> swizzled smem -> TMA store,
//            plus per-(sample, channel-quad) sum / sum-of-squares of the rounded outputs for the next GroupNorm.
//
// Replaces aten::conv2d (cuDNN) for every 3x3 / 1x1 conv with Cin % 64 == 0 and Cout % 128 == 0
// (reference call sites: diffusers ResnetBlock2D / Downsample2D / Upsample2D / Attention linears, built at
// /root/reference/train.py:788-810).
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "tensormap.h"

// Developer experiments (skip loads / skip the epilogue / per-role cycle counters) exist only in -DD3_DEV builds
// (D3_DEV=1 python -m d3roma_b200.build); the product library has no knob that can produce wrong results.
#ifdef D3_DEV
#define D3_DBG(a, mask) ((a).debug_skip & (mask))
#else
#define D3_DBG(a, mask) 0
#endif

namespace d3 {

namespace {

constexpr int kThreads = 256;
constexpr int kTileM = 128;
constexpr uint32_t kATileBytes = kTileM * 64 * 2;  // 16 KB
constexpr uint32_t kStagingBytes = kTileM * 128;   // one 64-channel chunk of the output tile

struct ConvKArgs {
  int B, Ho, Wo, Cout;
  int CoutV;               // channels that exist in the output tensor (<= Cout; the rest are weight padding)
  int TH, TW, tilesH, tilesW, nTilesN, numTiles;
  int splitK, kbPer;       // pixels-as-M kernel: K-blocks split over splitK CTAs per tile (fp32 partials in ws)
  float* ws;               // [splitK][B*Ho*Wo][Cout] fp32
  int ksize, stride, pad;
  int ntaps;
  signed char tap_dh[9], tap_dw[9];  // input offset of each filter tap (stride-1 path)
  int Cin, cbMain;
  int kbMain, kbSc1, kbSc2, kbTotal;
  int sc_off2;  // channel offset of the second half-res parity plane (stride 2 only): x_pitch
  const float* bias;
  int bias_bstride;
  const uint16_t* residual;
  int res_pitch;
  double* stats;
  uint32_t a_bytes;
  int blk_slot;    // halo-block variant: bytes per block ring slot ((TH+2)*TW*128)
  int n_blk, n_wts;  // halo-block variant: ring depths (<= kMaxBlk, kMaxWts)
  int pass_h;      // pixels-as-N kernel: tile rows per epilogue pass (1 unless TW < 64)
  int pass_w;      // pixels-as-N kernel: pixels per epilogue pass (divides TW, <= 64)
  int debug_skip;  // D3_DEV builds only: bit0 = skip A loads, bit1 = skip B loads, ... (results are garbage)
};

struct TileCoord {
  int b, h0, w0, n0;
};

#ifdef D3_DEV
// developer profiling (D3_CONV_DEBUG_SKIP bit 256): clock cycles summed over CTAs
//   [0] MMA thread total, [1] waiting for a free accumulator, [2] waiting for pixel blocks, [3] waiting for weights,
//   [4] epilogue thread total, [5] waiting for a full accumulator, [6] inside passes, [7] tiles
//   [8] pass: store-buffer wait + barrier 1, [9] TMEM load, [10] math + staging stores, [11] fence + barrier 2 + store issue
__device__ unsigned long long g_conv_prof[16];
#endif

D3_DEVINL TileCoord decode_tile(const ConvKArgs& a, int tile, int BN) {
  TileCoord t;
  int nt = tile % a.nTilesN;
  int p = tile / a.nTilesN;
  int tw = p % a.tilesW;
  p /= a.tilesW;
  int th = p % a.tilesH;
  t.b = p / a.tilesH;
  t.h0 = th * a.TH;
  t.w0 = tw * a.TW;
  t.n0 = nt * BN;
  return t;
}

// Butterfly transpose-reduce: on entry every lane holds 32 partial values v[0..31]; on exit lane l holds
// sum over all 32 lanes of the original v[l]. 31 shuffles.
D3_DEVINL float warp_transpose_reduce32(float (&v)[32], uint32_t lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = upper ? v[i] : v[i + off];
      float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}


// Epilogue math for 32 accumulator columns (half of a 64-channel chunk): + bias (+ residual), round to bf16,
// write the swizzled staging row, and collect per-quad sum / sum-of-squares of the rounded values.
template <int HALF, int DT>
D3_DEVINL void epi_process32(const uint32_t (&v)[32], float (&sums)[32], const float4* bp, const uint4* rp,
                             bool valid, uint8_t* sbuf, uint32_t r) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {  // 8 channels per iteration -> one 16 B smem store
    const int j8 = HALF * 4 + jj;
    float x[8];
    const float4 b0 = __ldg(bp + 2 * j8), b1 = __ldg(bp + 2 * j8 + 1);
    x[0] = __uint_as_float(v[8 * jj + 0]) + b0.x;
    x[1] = __uint_as_float(v[8 * jj + 1]) + b0.y;
    x[2] = __uint_as_float(v[8 * jj + 2]) + b0.z;
    x[3] = __uint_as_float(v[8 * jj + 3]) + b0.w;
    x[4] = __uint_as_float(v[8 * jj + 4]) + b1.x;
    x[5] = __uint_as_float(v[8 * jj + 5]) + b1.y;
    x[6] = __uint_as_float(v[8 * jj + 6]) + b1.z;
    x[7] = __uint_as_float(v[8 * jj + 7]) + b1.w;
    if (rp != nullptr) {
      const uint4 rv = __ldg(rp + j8);
      float2 f;
      f = unpack2<DT>(rv.x); x[0] += f.x; x[1] += f.y;
      f = unpack2<DT>(rv.y); x[2] += f.x; x[3] += f.y;
      f = unpack2<DT>(rv.z); x[4] += f.x; x[5] += f.y;
      f = unpack2<DT>(rv.w); x[6] += f.x; x[7] += f.y;
    }
    uint4 pk;
    pk.x = pack2<DT>(x[0], x[1]);
    pk.y = pack2<DT>(x[2], x[3]);
    pk.z = pack2<DT>(x[4], x[5]);
    pk.w = pack2<DT>(x[6], x[7]);
    // statistics of the values as stored (bf16-rounded), masked to in-image pixels
    const float2 f0 = unpack2<DT>(pk.x), f1 = unpack2<DT>(pk.y);
    const float2 f2 = unpack2<DT>(pk.z), f3 = unpack2<DT>(pk.w);
    const float s0 = (f0.x + f0.y) + (f1.x + f1.y);
    const float q0 = (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
    const float s1 = (f2.x + f2.y) + (f3.x + f3.y);
    const float q1 = (f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y);
    sums[2 * j8] = valid ? s0 : 0.f;
    sums[2 * j8 + 1] = valid ? s1 : 0.f;
    sums[16 + 2 * j8] = valid ? q0 : 0.f;
    sums[16 + 2 * j8 + 1] = valid ? q1 : 0.f;
    // SWIZZLE_128B: 16-byte chunk index XOR (row % 8)
    *reinterpret_cast<uint4*>(sbuf + r * 128 + ((j8 ^ (r & 7)) << 4)) = pk;
  }
}

template <int BN, int STAGES, int DT>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmS1,
                  const __grid_constant__ CUtensorMap tmS2, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const ConvKArgs a) {
  constexpr uint32_t kBBytes = BN * 128;
  constexpr uint32_t kStageBytes = kATileBytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512 (power of two)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* spart = reinterpret_cast<float*>(tmem_ptr + 4);  // [4][32]

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (a.kbSc1) tma_prefetch_desc(&tmS1);
    if (a.kbSc2) tma_prefetch_desc(&tmS2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_prologue();   // barriers / TMEM are set up; from here on the kernel reads what its predecessor wrote

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < a.numTiles * a.splitK; work += gridDim.x) {
        const int tile = work / a.splitK;
        const int kb0 = (work - tile * a.splitK) * a.kbPer;
        const int kb1 = min(a.kbTotal, kb0 + a.kbPer);
        const TileCoord t = decode_tile(a, tile, BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kATileBytes;
          if (D3_DBG(a, 3)) {
            const uint32_t bytes = (D3_DBG(a, 1) ? 0u : a.a_bytes) + (D3_DBG(a, 2) ? 0u : kBBytes);
            if (bytes) mbar_arrive_expect_tx(&full_bar[stage], bytes); else mbar_arrive(&full_bar[stage]);
            if (!D3_DBG(a, 1)) tma_load_4d(sA, &tmA, &full_bar[stage], 0, t.w0, t.h0, t.b);
            if (!D3_DBG(a, 2)) tma_load_2d(sB, &tmB, &full_bar[stage], kb * 64, t.n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], a.a_bytes + kBBytes);
          if (kb < a.kbMain) {
            const int tap = kb / a.cbMain;
            const int cb = kb - tap * a.cbMain;
            const int kh = tap / a.ksize;
            const int kw = tap - kh * a.ksize;
            if (a.stride == 1) {
              tma_load_4d(sA, &tmA, &full_bar[stage], cb * 64, t.w0 + a.tap_dw[tap], t.h0 + a.tap_dh[tap], t.b);
            } else {
              // input pixel (2*ho + kh - pad, 2*wo + kw - pad) in the parity-split 5-D view (pad = 1, or 0 for the
              // VAE encoder's F.pad(0,1,0,1) + unpadded stride-2 conv): row 2*hh + ph with hh = ho + floor(dh / 2)
              const int dh = kh - a.pad, dw = kw - a.pad;
              const int hh = t.h0 + (dh >> 1), ph = dh & 1;
              const int ww = t.w0 + (dw >> 1), pw = dw & 1;
              tma_load_5d(sA, &tmA, &full_bar[stage], pw * a.sc_off2 + cb * 64, ww, ph, hh, t.b);
            }
          } else if (kb < a.kbMain + a.kbSc1) {
            tma_load_4d(sA, &tmS1, &full_bar[stage], (kb - a.kbMain) * 64, t.w0, t.h0, t.b);
          } else {
            tma_load_4d(sA, &tmS2, &full_bar[stage], (kb - a.kbMain - a.kbSc1) * 64, t.w0, t.h0, t.b);
          }
          tma_load_2d(sB, &tmB, &full_bar[stage], kb * 64, t.n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_16bit(kTileM, BN, DT == kBF16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = blockIdx.x; work < a.numTiles * a.splitK; work += gridDim.x) {
        const int kb0 = (work % a.splitK) * a.kbPer;
        const int kb1 = min(a.kbTotal, kb0 + a.kbPer);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sA);
          const uint64_t bdesc = make_sw128_kmajor_desc(sA + kATileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const uint32_t q = warp - 4;
    const uint32_t et = threadIdx.x - 128;
    const uint32_t r = q * 32 + lane;  // accumulator row == pixel index inside the tile
    const int th = r / a.TW;
    const int tw = r - th * a.TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    for (int work = blockIdx.x; work < a.numTiles * a.splitK; work += gridDim.x) {
      const int tile = work / a.splitK;
      const TileCoord t = decode_tile(a, tile, BN);
      const int h = t.h0 + th, w = t.w0 + tw;
      const bool valid = (th < a.TH) && (h < a.Ho) && (w < a.Wo);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (a.splitK > 1) {
        // split-K: this CTA holds a partial sum over its K range - park it as fp32 [split][pixel][Cout]; the finalize
        // kernel adds the splits, bias and residual, rounds, stores and takes the statistics
        const int split = work - tile * a.splitK;
        float* dst = a.ws + ((size_t)split * a.B * a.Ho * a.Wo + ((size_t)(t.b * a.Ho + h) * a.Wo + w)) * a.Cout + t.n0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * BN + c * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(dst + c * 32 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (D3_DBG(a, 4)) {  // developer experiment: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t v0[32], v1[32];
        {
          const uint32_t taddr = tmem_base + ((q * 32) << 16) + acc * BN + c * 64;
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_ld_32x32b_x32(taddr + 32, v1);
          tmem_ld_wait();
        }
        if (c == BN / 64 - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        uint8_t* sbuf = staging + (chunk_ctr & 1) * kStagingBytes;
        if (et == 0) tma_store_wait_read<1>();  // the store that last used this buffer has drained
        named_bar_sync(1, 128);

        const int n_base = t.n0 + c * 64;
        const float4* bp = reinterpret_cast<const float4*>(a.bias + (size_t)t.b * a.bias_bstride + n_base);
        const uint4* rp = nullptr;
        if (a.residual != nullptr && valid && n_base < a.CoutV) {   // padded channel chunks have no residual
          rp = reinterpret_cast<const uint4*>(a.residual + ((size_t)(t.b * a.Ho + h) * a.Wo + w) * a.res_pitch +
                                              n_base);
        }
        float sums[32];  // [0..15] quad sums, [16..31] quad sums of squares
        epi_process32<0, DT>(v0, sums, bp, rp, valid, sbuf, r);
        epi_process32<1, DT>(v1, sums, bp, rp, valid, sbuf, r);
        if (a.stats != nullptr) {
          const float tot = warp_transpose_reduce32(sums, lane);
          spart[q * 32 + lane] = tot;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (et == 0) {
          if (n_base < a.CoutV) tma_store_4d(&tmO, sbuf, n_base, t.w0, t.h0, t.b);
          tma_store_commit();
        }
        if (a.stats != nullptr && et < 32) {
          const float tot = spart[et] + spart[32 + et] + spart[64 + et] + spart[96 + et];
          const int quad = (n_base >> 2) + (et & 15);
          atomicAdd(a.stats + ((size_t)t.b * (a.Cout >> 2) + quad) * 2 + (et >> 4), static_cast<double>(tot));
        }
        ++chunk_ctr;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES>
constexpr size_t conv_smem_bytes() {
  return 1024 + STAGES * (kATileBytes + BN * 128) + 2 * kStagingBytes + (2 * STAGES + 4) * 8 + 16 + 4 * 32 * 4;
}


// =====================================================================================================================
// "Pixels-as-N" variant for Cout tiles of exactly 128 channels (54 % of the UNet's FLOPs: every 128-channel layer).
//
// A single-CTA tcgen05.mma with N = 128 has to fetch 8 KB of smem operands every 64 clk (128 B/clk) and measures at
// ~50 % of the N = 256 rate on B200 (profiles/r01_conv_experiments.txt).  Swapping the operand roles puts the WEIGHTS
// on the M side (M = 128 output channels) and up to 256 PIXELS on the N side, so each instruction moves 12 KB per
// 128 clk (96 B/clk) like the BN = 256 kernel.  The accumulator is then D[channel lane][pixel column]; the epilogue
// transposes through the swizzled staging tile (2-byte stores, conflict-free) before the TMA store, and the per-channel
// GroupNorm partial sums become per-thread running sums (no shuffles except the final quad combine).
// =====================================================================================================================
constexpr uint32_t kWTileBytes = 128 * 128;          // 128 output channels x 64 k
constexpr uint32_t kPTileBytes = 256 * 128;          // up to 256 pixels x 64 k
constexpr uint32_t kTStageBytes = kWTileBytes + kPTileBytes;
constexpr uint32_t kTStagingBytes = 2 * kStagingBytes;  // 128 pixels x 128 channels (two 64-channel halves)


// Packed fp32x2 accumulate (Blackwell FADD2 / FFMA2): acc += x, sq += x * x on two lanes of a register pair.
D3_DEVINL void stats_acc2(float2& s, float2& q, float2 x) {
  asm("{\n\t.reg .b64 s, q, x;\n\t"
      "mov.b64 s, {%0, %1};\n\tmov.b64 q, {%2, %3};\n\tmov.b64 x, {%4, %5};\n\t"
      "add.rn.f32x2 s, s, x;\n\tfma.rn.f32x2 q, x, x, q;\n\t"
      "mov.b64 {%0, %1}, s;\n\tmov.b64 {%2, %3}, q;\n\t}"
      : "+f"(s.x), "+f"(s.y), "+f"(q.x), "+f"(q.y)
      : "f"(x.x), "f"(x.y));
}

D3_DEVINL float2 f2_add(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 a, b, r;\n\t"
      "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 r, a, b;\n\t"
      "mov.b64 {%0, %1}, r;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

// Lean chunk epilogue (per-tap kernel): each lane rounds ITS OWN channel at pixels J0 and J0+4 into one packed word, swaps it
// with the neighbouring channel's lane and assembles the 32-bit staging word with one PRMT (even lane: {ch, ch+1} at
// pixel J0; odd lane: {ch-1, ch} at pixel J0+4 - the same conflict-free store pattern as epi_t_chunk2, without its
// selects).  Statistics are per own channel on the rounded values, accumulated with packed fp32x2 ops; the MASKED
// instantiation (tiles that overhang the image) zeroes out-of-image pixels first.
template <int DT, bool HAS_RES, bool MASKED, int BASE = 0, int JJ = 0, bool NOSHFL = false>
D3_DEVINL void epi_t_chunk3(const uint32_t (&v)[32], uint32_t sa_lane, const uint32_t (&myswz)[4], uint32_t sa_own,
                            const uint32_t (&swz)[8], float bias, uint32_t vmask, uint32_t prmt_sel, float2& s2,
                            float2& q2) {
  if constexpr (BASE < 32) {
    constexpr int J0 = BASE + JJ, J1 = J0 + 4;
    float x0 = __uint_as_float(v[J0]) + bias;
    float x1 = __uint_as_float(v[J1]) + bias;
    if constexpr (HAS_RES) {
      x0 += unpack2<DT>(lds_u16<J0 * 128>(sa_own + swz[J0 & 7])).x;
      x1 += unpack2<DT>(lds_u16<J1 * 128>(sa_own + swz[J1 & 7])).x;
    }
    const uint32_t pk = pack2<DT>(x0, x1);
    const uint32_t other = NOSHFL ? pk : __shfl_xor_sync(0xffffffffu, pk, 1);
    sts_u32<J0 * 128>(sa_lane + myswz[JJ], __byte_perm(pk, other, prmt_sel));
    float2 xr = unpack2<DT>(pk);
    if constexpr (MASKED) {
      xr.x = ((vmask >> J0) & 1u) ? xr.x : 0.f;
      xr.y = ((vmask >> J1) & 1u) ? xr.y : 0.f;
    }
    stats_acc2(s2, q2, xr);
    if constexpr (JJ < 3)
      epi_t_chunk3<DT, HAS_RES, MASKED, BASE, JJ + 1, NOSHFL>(v, sa_lane, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
    else
      epi_t_chunk3<DT, HAS_RES, MASKED, BASE + 8, 0, NOSHFL>(v, sa_lane, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
  }
}

constexpr int kTThreads = 384;  // warps 0-3: producer / MMA / TMEM alloc / weight producer; warps 4-11: two epilogue groups

constexpr int kMaxBlk = 3;    // halo-block ring slots (halo-block variant)
constexpr int kMaxWts = 8;    // weight-tile ring slots (halo-block variant)

// ---------------------------------------------------------------------------------------------------------------------
// Transposing epilogue through stmatrix (halo-block kernel).  The accumulator is D[channel lane][pixel column]; the
// output tile in shared memory is [pixel row][channel] (what the NHWC TMA store needs).  tcgen05.ld.16x256b hands every
// warp 16 channels x 8 pixels per repetition in the mma C-fragment layout (thread T: channel T/4 (+8), pixels 2(T%4),
// 2(T%4)+1 - verified on B200 by tools/probes/probe_tmem_layout.cu), which is exactly the source fragment of
// stmatrix.m8n8.trans: after bias + rounding, ONE stmatrix.x4.trans writes four 8x8 blocks transposed as 16-byte
// (8-channel) row segments, conflict-free in the SWIZZLE_128B staging tile.  Per 32 x 32 chunk and warp: 2 TMEM loads,
// 4 stmatrix - against 16 shuffles + 16 PRMT + 16 STS.32 of the lane-exchange epilogue.  The residual (when given) is
// read back from the TMA-loaded staging tile with the mirror ldmatrix.x4.trans.  GroupNorm statistics are taken from
// the fp32 values before rounding (four channels per thread, packed f32x2 adds).
// ---------------------------------------------------------------------------------------------------------------------
D3_DEVINL void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
D3_DEVINL void stmatrix_x4_trans(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b),
               "r"(c), "r"(d)
               : "memory");
}
D3_DEVINL void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}

// 16 channels (one TMEM half-quarter) x 32 pixels: v = the 16 registers of one tcgen05.ld.16x256b.x4.
//   v[4i + 2hl + e] = D[channel base + T/4 + 8 hl][pixel 8i + 2(T%4) + e]
// sa = this thread's stmatrix row address for pixel block i = 0 (blocks 2, 3 are 2048 B further down).
template <int DT, bool HAS_RES, bool MASKED>
D3_DEVINL void epi_t_stsm16(const uint32_t (&v)[16], uint32_t sa, float bias_lo, float bias_hi, uint32_t vmask,
                            uint32_t t4, float2& s_lo, float2& q_lo, float2& s_hi, float2& q_hi) {
#pragma unroll
  for (int ii = 0; ii < 4; ii += 2) {
    uint32_t pk[4];
    uint32_t res[4];
    if constexpr (HAS_RES) ldmatrix_x4_trans(sa + ii * 1024, res);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = ii + (m >> 1), hl = m & 1;
      const float bsel = hl ? bias_hi : bias_lo;
      float2 x = f2_add(make_float2(__uint_as_float(v[4 * i + 2 * hl]), __uint_as_float(v[4 * i + 2 * hl + 1])),
                        make_float2(bsel, bsel));
      if constexpr (HAS_RES) x = f2_add(x, unpack2<DT>(res[m]));
      pk[m] = pack2<DT>(x.x, x.y);
      if constexpr (MASKED) {
        const uint32_t bit = 8u * i + 2u * t4;
        x.x = ((vmask >> bit) & 1u) ? x.x : 0.f;
        x.y = ((vmask >> (bit + 1)) & 1u) ? x.y : 0.f;
      }
      if (hl) stats_acc2(s_hi, q_hi, x); else stats_acc2(s_lo, q_lo, x);
    }
    stmatrix_x4_trans(sa + ii * 1024, pk[0], pk[1], pk[2], pk[3]);
  }
}

// STRIPS = true ("halo blocks"): 3x3 stride-1 convs.  The pixel operand of a TH x TW tile (TW a multiple of 8, so one
// tile row = TW*128 B is a whole number of 1024 B swizzle atoms) is loaded ONCE per (kw, 64-channel block) as a block of
// TH+2 rows starting one row above the tile, and the three kh taps read it through UMMA descriptors whose start
// address is advanced by whole tile rows - atom-aligned, unlike a kw (single-pixel) shift, which measured ~25 % slower
// in the MMA itself (profiles/r01_conv_experiments.txt).  Activation TMA traffic drops by 3*TH/(TH+2).  Weights stream
// through their own ring, fed by a second producer lane (warp 3).
// K-loop order: for cb: for kw: block(kw, cb) -> for kh: weight tile (kh, kw, cb) -> 4 MMAs.
// EPI: 0 = lane-exchange epilogue (epi_t_chunk3), 1 = stmatrix epilogue (epi_t_stsm16; needs pass windows that start on
// multiples of 8 accumulator columns, which the halo-block tiles guarantee).
template <int STAGES, int DT, bool STRIPS, int EPI>
__global__ void __launch_bounds__(kTThreads, 1)
conv_igemm_t_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmS1,
                    const __grid_constant__ CUtensorMap tmS2, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                    const ConvKArgs a) {
  constexpr uint32_t kTmemCols = 512;  // 2 accumulator stages x 256 pixel columns
  constexpr int NBAR = STRIPS ? (kMaxBlk + kMaxWts) : STAGES;  // ring barriers (full + empty each)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // STRIPS: [n_blk block slots of a.blk_slot bytes][n_wts weight tiles of 16 KB]; else STAGES x {16 KB W, 32 KB P}
  uint8_t* wring = STRIPS ? smem + a.n_blk * a.blk_slot : smem;
  uint8_t* staging = STRIPS ? wring + a.n_wts * kWTileBytes : smem + STAGES * kTStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kTStagingBytes);
  uint64_t* empty_bar = full_bar + NBAR;
  uint64_t* tfull_bar = empty_bar + NBAR;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 1);
  // halo-block variant: barriers [0, kMaxBlk) belong to the block ring, [kMaxBlk, NBAR) to the weight ring
  uint64_t* sfull = full_bar;
  uint64_t* sempty = empty_bar;
  uint64_t* wfull = full_bar + kMaxBlk;
  uint64_t* wempty = empty_bar + kMaxBlk;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int NP = a.TH * a.TW;                 // pixels per tile (<= 256)
  const uint32_t n_mma = (NP + 15) & ~15;     // MMA N (multiple of 16)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NBAR; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_prologue();   // barriers / TMEM are set up; from here on the kernel reads what its predecessor wrote

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (STRIPS && lane == 0) {
      int ss = 0;
      uint32_t sphase = 0;
      const uint32_t blk_bytes = (uint32_t)(a.TH + 2) * a.TW * 128u;
      const uint32_t plain_bytes = (uint32_t)a.TH * a.TW * 128u;
      const int n_items = 3 * a.cbMain + a.kbSc1 + a.kbSc2;  // block loads per tile
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(a, tile, 128);
        for (int it = 0; it < n_items; ++it) {
          mbar_wait(&sempty[ss], sphase ^ 1);
          uint8_t* sS = smem + ss * a.blk_slot;
          if (D3_DBG(a, 1)) {  // developer experiment: no activation loads
            mbar_arrive(&sfull[ss]);
          } else if (it < 3 * a.cbMain) {
            const int cb = it / 3, kw = it - cb * 3;
            mbar_arrive_expect_tx(&sfull[ss], blk_bytes);
            tma_load_4d(sS, &tmA, &sfull[ss], cb * 64, t.w0 + kw - 1, t.h0 - 1, t.b);
          } else {
            const int j = it - 3 * a.cbMain;
            mbar_arrive_expect_tx(&sfull[ss], plain_bytes);
            if (j < a.kbSc1) tma_load_4d(sS, &tmS1, &sfull[ss], j * 64, t.w0, t.h0, t.b);
            else tma_load_4d(sS, &tmS2, &sfull[ss], (j - a.kbSc1) * 64, t.w0, t.h0, t.b);
          }
          if (++ss == a.n_blk) { ss = 0; sphase ^= 1; }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(a, tile, 128);
        for (int kb = 0; kb < a.kbTotal; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sW = smem + stage * kTStageBytes;
          uint8_t* sP = sW + kWTileBytes;
          if (D3_DBG(a, 3)) {
            const uint32_t bytes = (D3_DBG(a, 1) ? 0u : a.a_bytes) + (D3_DBG(a, 2) ? 0u : kWTileBytes);
            if (bytes) mbar_arrive_expect_tx(&full_bar[stage], bytes); else mbar_arrive(&full_bar[stage]);
            if (!D3_DBG(a, 1)) tma_load_4d(sP, &tmA, &full_bar[stage], 0, t.w0, t.h0, t.b);
            if (!D3_DBG(a, 2)) tma_load_2d(sW, &tmB, &full_bar[stage], kb * 64, t.n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], a.a_bytes + kWTileBytes);
          if (kb < a.kbMain) {
            const int tap = kb / a.cbMain;
            const int cb = kb - tap * a.cbMain;
            const int kh = tap / a.ksize;
            const int kw = tap - kh * a.ksize;
            if (a.stride == 1) {
              tma_load_4d(sP, &tmA, &full_bar[stage], cb * 64, t.w0 + a.tap_dw[tap], t.h0 + a.tap_dh[tap], t.b);
            } else {
              const int dh = kh - a.pad, dw = kw - a.pad;
              const int hh = t.h0 + (dh >> 1), ph = dh & 1;
              const int ww = t.w0 + (dw >> 1), pw = dw & 1;
              tma_load_5d(sP, &tmA, &full_bar[stage], pw * a.sc_off2 + cb * 64, ww, ph, hh, t.b);
            }
          } else if (kb < a.kbMain + a.kbSc1) {
            tma_load_4d(sP, &tmS1, &full_bar[stage], (kb - a.kbMain) * 64, t.w0, t.h0, t.b);
          } else {
            tma_load_4d(sP, &tmS2, &full_bar[stage], (kb - a.kbMain - a.kbSc1) * 64, t.w0, t.h0, t.b);
          }
          tma_load_2d(sW, &tmB, &full_bar[stage], kb * 64, t.n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (STRIPS && lane == 0) {
      const uint32_t idesc = make_idesc_16bit(kTileM, n_mma, DT == kBF16);
      int ss = 0, ws = 0;
      uint32_t sphase = 0, wphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t row_bytes = (uint32_t)a.TW * 128u;
      const int n_items = 3 * a.cbMain + a.kbSc1 + a.kbSc2;
#ifdef D3_DEV
      const bool prof = D3_DBG(a, 256) != 0;
      long long p_t0 = prof ? clock64() : 0, p_acc = 0, p_blk = 0, p_wts = 0, p_c = 0;
#define D3_PROF(stmt) do { if (prof) { stmt; } } while (0)
#else
#define D3_PROF(stmt) do { } while (0)
#endif
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        D3_PROF(p_c = clock64());
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        D3_PROF(p_acc += clock64() - p_c);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int it = 0; it < n_items; ++it) {
          D3_PROF(p_c = clock64());
          mbar_wait(&sfull[ss], sphase);
          D3_PROF(p_blk += clock64() - p_c);
          tc_fence_after();
          const uint32_t sS = smem_u32(smem + ss * a.blk_slot);
          const int nkb = (it < 3 * a.cbMain) ? 3 : 1;
          for (int i = 0; i < nkb; ++i) {
            D3_PROF(p_c = clock64());
            mbar_wait(&wfull[ws], wphase);
            D3_PROF(p_wts += clock64() - p_c);
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(wring + ws * kWTileBytes));
            // pixel window of tap kh = i: tile rows i .. i+TH-1 of the block (whole swizzle atoms further down)
            const uint64_t bdesc = make_sw128_kmajor_desc(sS + (uint32_t)i * row_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
            }
            first = 0;
            umma_commit(&wempty[ws]);
            if (++ws == a.n_wts) { ws = 0; wphase ^= 1; }
          }
          umma_commit(&sempty[ss]);
          if (++ss == a.n_blk) { ss = 0; sphase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
#ifdef D3_DEV
      if (prof) {
        atomicAdd(&g_conv_prof[0], (unsigned long long)(clock64() - p_t0));
        atomicAdd(&g_conv_prof[1], (unsigned long long)p_acc);
        atomicAdd(&g_conv_prof[2], (unsigned long long)p_blk);
        atomicAdd(&g_conv_prof[3], (unsigned long long)p_wts);
      }
#endif
#undef D3_PROF
    } else if (lane == 0) {
      const uint32_t idesc = make_idesc_16bit(kTileM, n_mma, DT == kBF16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < a.kbTotal; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sW = smem_u32(smem + stage * kTStageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sW);                 // weights: M = 128 channels
          const uint64_t bdesc = make_sw128_kmajor_desc(sW + kWTileBytes);   // pixels:  N = n_mma
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp == 3) {
    // ================================ weight producer (halo-block variant) ================================
    if (STRIPS && lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      const int n_items = 3 * a.cbMain + a.kbSc1 + a.kbSc2;
      for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(a, tile, 128);
        for (int it = 0; it < n_items; ++it) {
          int nkb = 1, kb0 = a.kbMain + (it - 3 * a.cbMain);
          if (it < 3 * a.cbMain) {
            const int cb = it / 3, kw = it - cb * 3;
            nkb = 3;
            kb0 = kw * a.cbMain + cb;  // tap (kh, kw) -> K-block (kh*3 + kw) * cbMain + cb
          }
          for (int i = 0; i < nkb; ++i) {
            mbar_wait(&wempty[ws], wphase ^ 1);
            if (D3_DBG(a, 2)) {
              mbar_arrive(&wfull[ws]);
            } else {
              mbar_arrive_expect_tx(&wfull[ws], kWTileBytes);
              tma_load_2d(wring + ws * kWTileBytes, &tmB, &wfull[ws], (kb0 + i * 3 * a.cbMain) * 64, t.n0);
            }
            if (++ws == a.n_wts) { ws = 0; wphase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const uint32_t q = (warp - 4) & 3;          // TMEM lane quarter this warp may access
    const uint32_t grp = (warp - 4) >> 2;       // the two groups take the two 32-pixel chunks of a pass
    const uint32_t et = threadIdx.x - 128;
    const bool has_res = a.residual != nullptr;
    // A pass = PW (<= 64) consecutive pixels of one tile row = a contiguous range of accumulator columns.  Staging is
    // double-buffered by pass so the TMA store of pass p overlaps the math of pass p+1.
    const int PW = a.pass_w;
    const int PH = a.pass_h;                    // > 1 only for narrow tiles (TW < 64): a pass = PH whole tile rows
    const int PN = PW * PH;                     // pixels (accumulator columns) per pass
    const int segs = a.TW / PW;
    const int passes = (a.TH / PH) * segs;
    // ---- lane-exchange epilogue (EPI == 0): one channel per thread
    const uint32_t ch = q * 32 + lane;          // channel inside the 128-channel tile == TMEM lane
    const uint32_t half = ch >> 6;              // which 64-channel staging half
    const uint32_t ch16 = (ch & 63) >> 3;       // 16-byte chunk of this channel inside the 128-byte row
    const uint32_t chb = (ch & 7) * 2;          // byte inside the chunk
    uint32_t swz[8];  // swizzled 16-byte chunk offset of this channel for row (pixel) index % 8
#pragma unroll
    for (int i = 0; i < 8; ++i) swz[i] = (ch16 ^ (uint32_t)i) << 4;
    const bool odd = (lane & 1u) != 0;
    uint32_t myswz[4];  // swizzle offsets of the 4 pixel rows (mod 8) this lane stores: even lanes 0-3, odd lanes 4-7
#pragma unroll
    for (int i = 0; i < 4; ++i) myswz[i] = swz[i + (odd ? 4 : 0)];
    const uint32_t prmt_sel = odd ? 0x3276u : 0x5410u;  // see epi_t_chunk3
    // ---- stmatrix epilogue (EPI == 1): four channels per thread: 32q + 16 ab + 8 hl + lane/4
    const uint32_t t4 = lane & 3, l4 = lane >> 2;
    const uint32_t sm_m = lane >> 3, sm_j = lane & 7;       // stmatrix: this lane addresses row j of matrix m
    uint32_t sm_off[2];                                     // per ab: row offset of this lane inside the staging buffer
#pragma unroll
    for (int ab = 0; ab < 2; ++ab) {
      const uint32_t cb16 = (q & 1u) * 4u + 2u * ab + (sm_m & 1u);   // 16-byte channel chunk inside the 64-channel half
      sm_off[ab] = (q >> 1) * (kStagingBytes / 2) + (8u * (sm_m >> 1) + sm_j) * 128u + ((cb16 ^ sm_j) << 4);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    uint32_t pass_ctr = 0;
#ifdef D3_DEV
    const bool prof = D3_DBG(a, 256) != 0 && et == 0;
    long long p_t0 = prof ? clock64() : 0, p_wait = 0, p_pass = 0, p_c = 0, p_tiles = 0;
    long long p_q[4] = {0, 0, 0, 0};
    long long q0 = 0, q1 = 0, q2c = 0, q3 = 0;
#define D3_PROF(stmt) do { if (prof) { stmt; } } while (0)
#else
#define D3_PROF(stmt) do { } while (0)
#endif
    for (int tile = blockIdx.x; tile < a.numTiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(a, tile, 128);
      const float* bias_p = a.bias + (size_t)t.b * a.bias_bstride + t.n0;
      const float bias = bias_p[ch];
      float s_sum = 0.f, s_sq = 0.f;                                  // EPI 0
      float b4[4];                                                    // EPI 1: bias of the thread's four channels
      float2 s4[4], q4[4];                                            //        and their packed partial statistics
      if constexpr (EPI == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          b4[k] = bias_p[q * 32 + 8 * k + l4];
          s4[k] = make_float2(0.f, 0.f);
          q4[k] = make_float2(0.f, 0.f);
        }
      }
      D3_PROF(p_c = clock64());
      mbar_wait(&tfull_bar[acc], acc_phase);
      D3_PROF(const long long n = clock64(); p_wait += n - p_c; p_c = n; ++p_tiles);
      tc_fence_after();
      if (D3_DBG(a, 4)) {  // developer experiment: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      for (int ps = 0; ps < passes; ++ps, ++pass_ctr) {
        const int r = (ps / segs) * PH;
        const int sg = ps - (ps / segs) * segs;
        const int h = t.h0 + r;
        const int wpass = t.w0 + sg * PW;
        uint8_t* sbuf = staging + (pass_ctr & 1) * kStagingBytes;  // two 64-channel halves of 64 rows each
        D3_PROF(q0 = clock64());
        if (et == 0) {
          tma_store_wait_read<1>();  // the store that used this buffer two passes ago has finished reading it
          if (has_res) {
            mbar_arrive_expect_tx(res_bar, 2u * (uint32_t)PN * 128u);
            tma_load_4d(sbuf, &tmR, res_bar, t.n0, wpass, h, t.b);
            tma_load_4d(sbuf + kStagingBytes / 2, &tmR, res_bar, t.n0 + 64, wpass, h, t.b);
          }
        }
        named_bar_sync(1, 256);
        if (has_res) {
          mbar_wait(res_bar, res_phase);
          res_phase ^= 1;
        }
        const int c0 = (int)grp * 32;
        D3_PROF(q1 = clock64());
        if (c0 < PN) {
          const int pc = c0 + (int)lane;          // pixel of the pass; row pc / PW, column pc % PW (PH == 1: row 0)
          const int prow = (PH > 1) ? pc / PW : 0;
          const int pcol = pc - prow * PW;
          const uint32_t vmask = __ballot_sync(0xffffffffu, (pc < PN) && (h + prow < a.Ho) && (wpass + pcol < a.Wo));
          const bool full = (vmask == 0xffffffffu);
          const uint32_t tcol = tmem_base + ((q * 32) << 16) + acc * 256 + r * a.TW + sg * PW + c0;
          if constexpr (EPI == 1) {
            uint32_t va[16], vb[16];
            tmem_ld_16x256b_x4(tcol, va);
            tmem_ld_16x256b_x4(tcol + (16u << 16), vb);
            const uint32_t sa = smem_u32(sbuf) + (uint32_t)c0 * 128u;
            tmem_ld_wait();
            D3_PROF(q2c = clock64());
            if (D3_DBG(a, 32)) {  // developer experiment: no epilogue math
            } else if (has_res) {
              if (full) {
                epi_t_stsm16<DT, true, false>(va, sa + sm_off[0], b4[0], b4[1], vmask, t4, s4[0], q4[0], s4[1], q4[1]);
                epi_t_stsm16<DT, true, false>(vb, sa + sm_off[1], b4[2], b4[3], vmask, t4, s4[2], q4[2], s4[3], q4[3]);
              } else {
                epi_t_stsm16<DT, true, true>(va, sa + sm_off[0], b4[0], b4[1], vmask, t4, s4[0], q4[0], s4[1], q4[1]);
                epi_t_stsm16<DT, true, true>(vb, sa + sm_off[1], b4[2], b4[3], vmask, t4, s4[2], q4[2], s4[3], q4[3]);
              }
            } else {
              if (full) {
                epi_t_stsm16<DT, false, false>(va, sa + sm_off[0], b4[0], b4[1], vmask, t4, s4[0], q4[0], s4[1], q4[1]);
                epi_t_stsm16<DT, false, false>(vb, sa + sm_off[1], b4[2], b4[3], vmask, t4, s4[2], q4[2], s4[3], q4[3]);
              } else {
                epi_t_stsm16<DT, false, true>(va, sa + sm_off[0], b4[0], b4[1], vmask, t4, s4[0], q4[0], s4[1], q4[1]);
                epi_t_stsm16<DT, false, true>(vb, sa + sm_off[1], b4[2], b4[3], vmask, t4, s4[2], q4[2], s4[3], q4[3]);
              }
            }
          } else {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tcol, v);
            // word address of this lane's channel pair; odd lanes write 4 pixel rows further down (see epi_t_chunk3)
            const uint32_t sa = smem_u32(sbuf) + half * (kStagingBytes / 2) + (uint32_t)c0 * 128u + (chb & ~3u) +
                                (odd ? 512u : 0u);
            // byte address of this lane's own channel at pixel row 0 of the chunk (residual read-back)
            const uint32_t sa_own = smem_u32(sbuf) + half * (kStagingBytes / 2) + (uint32_t)c0 * 128u + chb;
            tmem_ld_wait();
            D3_PROF(q2c = clock64());
            float2 s2 = make_float2(s_sum, 0.f), q2 = make_float2(s_sq, 0.f);
            if (D3_DBG(a, 32)) {  // developer experiment: no epilogue math
            } else if (has_res) {
              if (full) epi_t_chunk3<DT, true, false>(v, sa, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
              else epi_t_chunk3<DT, true, true>(v, sa, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
            } else {
              if (full) epi_t_chunk3<DT, false, false>(v, sa, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
              else epi_t_chunk3<DT, false, true>(v, sa, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
            }
            s_sum = s2.x + s2.y;
            s_sq = q2.x + q2.y;
          }
        }
        D3_PROF(q3 = clock64());
        if (ps == passes - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (et == 0 && !D3_DBG(a, 8)) {
          if (t.n0 < a.CoutV) tma_store_4d(&tmO, sbuf, t.n0, wpass, h, t.b);
          if (t.n0 + 64 < a.CoutV) tma_store_4d(&tmO, sbuf + kStagingBytes / 2, t.n0 + 64, wpass, h, t.b);
          tma_store_commit();
        }
        D3_PROF(const long long q4c = clock64(); p_q[0] += q1 - q0; p_q[1] += q2c - q1; p_q[2] += q3 - q2c;
                p_q[3] += q4c - q3);
      }
      D3_PROF(p_pass += clock64() - p_c);
      if (a.stats != nullptr) {
        if constexpr (EPI == 1) {
          // thread holds, for channels 32q + 8k + lane/4 (k = 0..3), partial sums over its pixel columns: reduce over
          // the 4 lanes of a channel (xor 1, 2), then over the 4 channels of a quad (lane/4 = 4m..4m+3: xor 4, 8)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float s = s4[k].x + s4[k].y, qq = q4[k].x + q4[k].y;
#pragma unroll
            for (int off = 1; off <= 8; off <<= 1) {
              s += __shfl_xor_sync(0xffffffffu, s, off);
              qq += __shfl_xor_sync(0xffffffffu, qq, off);
            }
            if ((lane & 15u) == 0) {   // lane 0: channels 8k .. 8k+3 of the quarter; lane 16: channels 8k+4 .. 8k+7
              double* dst = a.stats + ((size_t)t.b * (a.Cout >> 2) + ((t.n0 + q * 32 + 8 * k) >> 2) + (lane >> 4)) * 2;
              atomicAdd(dst, (double)s);
              atomicAdd(dst + 1, (double)qq);
            }
          }
        } else {
          // quad = 4 consecutive channels = 4 consecutive lanes
          s_sum += __shfl_xor_sync(0xffffffffu, s_sum, 1);
          s_sq += __shfl_xor_sync(0xffffffffu, s_sq, 1);
          s_sum += __shfl_xor_sync(0xffffffffu, s_sum, 2);
          s_sq += __shfl_xor_sync(0xffffffffu, s_sq, 2);
          if ((lane & 3) == 0) {
            double* dst = a.stats + ((size_t)t.b * (a.Cout >> 2) + ((t.n0 + ch) >> 2)) * 2;
            atomicAdd(dst, (double)s_sum);
            atomicAdd(dst + 1, (double)s_sq);
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
#ifdef D3_DEV
    if (prof) {
      atomicAdd(&g_conv_prof[4], (unsigned long long)(clock64() - p_t0));
      atomicAdd(&g_conv_prof[5], (unsigned long long)p_wait);
      atomicAdd(&g_conv_prof[6], (unsigned long long)p_pass);
      atomicAdd(&g_conv_prof[7], (unsigned long long)p_tiles);
      for (int i = 0; i < 4; ++i) atomicAdd(&g_conv_prof[8 + i], (unsigned long long)p_q[i]);
    }
#endif
#undef D3_PROF
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int STAGES>
constexpr size_t conv_t_smem_bytes() {
  return 1024 + STAGES * kTStageBytes + kTStagingBytes + (2 * STAGES + 5) * 8 + 16;
}
inline size_t conv_ts_smem_bytes(const ConvKArgs& a) {
  return 1024 + (size_t)a.n_blk * a.blk_slot + (size_t)a.n_wts * kWTileBytes + kTStagingBytes +
         (2 * (kMaxBlk + kMaxWts) + 5) * 8 + 16;
}

template <int STAGES, int DT>
cudaError_t launch_conv_t(int grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmS1,
                          const CUtensorMap& tmS2, const CUtensorMap& tmB, const CUtensorMap& tmO,
                          const CUtensorMap& tmR, const ConvKArgs& a, bool strips, int epi) {
  if (strips) {
    const size_t smem = conv_ts_smem_bytes(a);
    if (epi == 1) {
      cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(conv_igemm_t_kernel<STAGES, DT, true, 1>), smem);
      if (e != cudaSuccess) return e;
      return launch_pdl(conv_igemm_t_kernel<STAGES, DT, true, 1>, dim3(grid), dim3(kTThreads), smem, stream, tmA, tmS1,
                        tmS2, tmB, tmO, tmR, a);
    } else {
      cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(conv_igemm_t_kernel<STAGES, DT, true, 0>), smem);
      if (e != cudaSuccess) return e;
      return launch_pdl(conv_igemm_t_kernel<STAGES, DT, true, 0>, dim3(grid), dim3(kTThreads), smem, stream, tmA, tmS1,
                        tmS2, tmB, tmO, tmR, a);
    }
    return cudaSuccess;
  }
  constexpr size_t smem = conv_t_smem_bytes<STAGES>();
  cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(conv_igemm_t_kernel<STAGES, DT, false, 0>), smem);
  if (e != cudaSuccess) return e;
  return launch_pdl(conv_igemm_t_kernel<STAGES, DT, false, 0>, dim3(grid), dim3(kTThreads), smem, stream, tmA, tmS1,
                    tmS2, tmB, tmO, tmR, a);
}

// Pixel tile for the pixels-as-N kernel: TH*TW <= 256; the epilogue walks it in passes of PW pixels of one row,
// PW = largest divisor of TW that is <= 64.
int pass_width(int tw) {
  for (int d = (tw < 64 ? tw : 64); d >= 1; --d)
    if (tw % d == 0) return d;
  return 1;
}

// The epilogue reads the accumulator in 32-column windows starting at r*TW + sg*PW + {0, 32}: every window must end
// inside the 256-column accumulator stage (a window running past column 512 of the second stage is a TMEM fault).
bool t_tile_windows_ok(int th, int tw) {
  const int pw = pass_width(tw);
  const int segs = tw / pw;
  const int c0max = (pw > 32) ? 32 : 0;
  return (th - 1) * tw + (segs - 1) * pw + c0max + 32 <= 256;
}

void choose_tile_t(int Ho, int Wo, int* TH, int* TW) {
  double best = -1.0;
  int bh = 1, bw = 1;
  for (int tw = 1; tw <= 256 && tw <= Wo; ++tw) {
    int th = 256 / tw;
    if (th > Ho) th = Ho;
    while (th > 1 && !t_tile_windows_ok(th, tw)) --th;
    if (th < 1 || !t_tile_windows_ok(th, tw)) continue;
    const int pw = pass_width(tw);
    if (pw < 16 && tw > 16) continue;  // avoid degenerate pass widths (e.g. prime tile widths)
    const double tiles = double(ceil_div(Ho, th)) * ceil_div(Wo, tw);
    double eff = double(Ho) * Wo / (tiles * 256.0);
    eff *= (pw >= 32 ? 1.0 : 0.98);    // prefer tiles whose passes fill both epilogue groups
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > bw)) {
      best = eff;
      bh = th;
      bw = tw;
    }
  }
  *TH = bh;
  *TW = bw;
}


// Split-K finalize: out = round(sum_s ws[s] + bias (+ residual)), GroupNorm quad statistics of the rounded values.
// Thread = one pixel slot x 8 channels; blockDim.x = (Cout / 8) * k.
struct SplitKFinArgs {
  const float* ws;
  int splitK;
  const float* bias;
  int bias_bstride;
  const uint16_t* residual;
  int res_pitch;
  uint16_t* out;
  int out_pitch;
  double* stats;
  int B, HW, Cout;
};

template <int DT>
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const SplitKFinArgs a) {
  pdl_prologue();
  const int nvec = a.Cout >> 3;
  const int vec = threadIdx.x % nvec, prow = threadIdx.x / nvec, pstride = blockDim.x / nvec;
  const int b = blockIdx.y;
  const size_t M = (size_t)a.B * a.HW;
  float bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias8[j] = a.bias[(size_t)b * a.bias_bstride + vec * 8 + j];
  float st[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p = blockIdx.x * pstride + prow; p < a.HW; p += gridDim.x * pstride) {
    const size_t pix = (size_t)b * a.HW + p;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias8[j];
    for (int s = 0; s < a.splitK; ++s) {
      const float4* src = reinterpret_cast<const float4*>(a.ws + ((size_t)s * M + pix) * a.Cout + vec * 8);
      const float4 u = __ldcs(src), v = __ldcs(src + 1);
      acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w;
      acc[4] += v.x; acc[5] += v.y; acc[6] += v.z; acc[7] += v.w;
    }
    if (a.residual != nullptr) {
      const uint4 rv = *reinterpret_cast<const uint4*>(a.residual + pix * a.res_pitch + vec * 8);
      float2 f;
      f = unpack2<DT>(rv.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack2<DT>(rv.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack2<DT>(rv.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack2<DT>(rv.w); acc[6] += f.x; acc[7] += f.y;
    }
    uint4 pk;
    pk.x = pack2<DT>(acc[0], acc[1]);
    pk.y = pack2<DT>(acc[2], acc[3]);
    pk.z = pack2<DT>(acc[4], acc[5]);
    pk.w = pack2<DT>(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(a.out + pix * a.out_pitch + vec * 8) = pk;
    const float2 f0 = unpack2<DT>(pk.x), f1 = unpack2<DT>(pk.y), f2 = unpack2<DT>(pk.z), f3 = unpack2<DT>(pk.w);
    st[0] += (f0.x + f0.y) + (f1.x + f1.y);
    st[1] += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
    st[2] += (f2.x + f2.y) + (f3.x + f3.y);
    st[3] += (f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y);
  }
  if (a.stats != nullptr) {
    // combine the pixel slots of the block in shared memory (fixed order), then ONE fp64 atomic per value and block -
    // per-thread atomics on the few hundred statistics addresses serialise in L2 (measured 40 us per launch)
    __shared__ float s_red[256 * 4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s_red[threadIdx.x * 4 + i] = st[i];
    __syncthreads();
    if (prow == 0) {
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      for (int r = 0; r < pstride; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] += (double)s_red[(r * nvec + vec) * 4 + i];
      double* dst = a.stats + ((size_t)b * (a.Cout >> 2) + 2 * vec) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dst + i, t[i]);
    }
  }
}

template <int BN, int STAGES, int DT>
cudaError_t launch_conv(int grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmS1,
                        const CUtensorMap& tmS2, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvKArgs& a) {
  constexpr size_t smem = conv_smem_bytes<BN, STAGES>();
  cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(conv_igemm_kernel<BN, STAGES, DT>), smem);
  if (e != cudaSuccess) return e;
  return launch_pdl(conv_igemm_kernel<BN, STAGES, DT>, dim3(grid), dim3(kThreads), smem, stream, tmA, tmS1, tmS2, tmB,
                    tmO, a);
}

// Pick the pixel tile TH x TW (TH*TW <= 128) that wastes the fewest MMA rows.
void choose_tile(int Ho, int Wo, int* TH, int* TW) {
  double best = -1.0;
  int bh = 1, bw = 1;
  for (int tw = 1; tw <= 128 && tw <= Wo; ++tw) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th < 1) continue;
    if (th > 256 || tw > 256) continue;
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

}  // namespace d3

namespace d3 {
namespace {
// Split-K plan of the pixels-as-M kernel: when a layer has too few tiles to occupy half the SMs (small feature maps,
// small batches) its K loop is cut into S ranges handled by S CTAs; 1 = no split.  Shared by d3_conv2d_igemm and
// d3_conv2d_workspace_bytes.  D3_CONV_SPLITK=0 disables.
int plan_split_k(const d3_conv_params* p, int* BN_out) {
  const int Ho = p->H / p->stride, Wo = p->W / p->stride;
  int TH, TW;
  choose_tile(Ho, Wo, &TH, &TW);
  const int pixTiles = p->B * ceil_div(Ho, TH) * ceil_div(Wo, TW);
  int BN = (p->Cout % 256 == 0) ? 256 : 128;
  if (BN == 256 && pixTiles * (p->Cout / 256) < num_sms()) BN = 128;
  if (BN_out) *BN_out = BN;
  static int want = -1;
  if (want < 0) { const char* e = getenv("D3_CONV_SPLITK"); want = e ? atoi(e) : 4; }
  if (!want || p->up2x_phase != 0 || p->Cout == 128) return 1;
  if (p->Cout_valid > 0 && p->Cout_valid != p->Cout) return 1;   // the finalize kernel writes all Cout channels
  if (p->Cout > 2048) return 1;                                   // ... with one thread per 8 channels of a pixel
  const int numTiles = pixTiles * (p->Cout / BN);
  const int kbTotal = (p->ksize * p->ksize * p->Cin + (p->sc1 ? p->C1 : 0) + (p->sc2 ? p->C2 : 0)) / 64;
  if (numTiles * 2 > num_sms() || kbTotal < 8) return 1;
  int S = num_sms() / numTiles;
  if (S > kbTotal / 8) S = kbTotal / 8;   // at least 8 K-blocks per CTA
  if (S > want) S = want;                 // D3_CONV_SPLITK = largest split factor (the fp32 partials cost S x the output)
  if (S < 2) return 1;
  const int per = ceil_div(kbTotal, S);
  return ceil_div(kbTotal, per);
}
}  // namespace
}  // namespace d3

extern "C" size_t d3_conv2d_workspace_bytes(const d3_conv_params* p) {
  using namespace d3;
  if (p == nullptr || p->dtype == D3_DTYPE_F32 || p->Cin <= 0 || p->Cin % 64 || p->Cout <= 0 || p->Cout % 128 ||
      (p->stride != 1 && p->stride != 2) || (p->ksize != 1 && p->ksize != 3))
    return 0;
  const int S = plan_split_k(p, nullptr);
  if (S <= 1) return 0;
  return (size_t)S * p->B * (p->H / p->stride) * (p->W / p->stride) * p->Cout * sizeof(float);
}

#ifdef D3_DEV
// Developer hook (D3_DEV builds only; not part of include/d3roma_b200.h): read (and optionally clear) the cycle
// counters collected while D3_CONV_DEBUG_SKIP has bit 256 set.
extern "C" int d3_debug_conv_profile(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, d3::g_conv_prof, sizeof(d3::g_conv_prof)) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(d3::g_conv_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif

extern "C" int d3_conv2d_igemm(const d3_conv_params* p, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(p != nullptr, "conv: null params");
  D3_REQUIRE(p->ksize == 3 || p->ksize == 1, "conv: ksize must be 1 or 3 (got %d)", p->ksize);
  D3_REQUIRE(p->stride == 1 || (p->stride == 2 && p->ksize == 3), "conv: unsupported stride %d", p->stride);
  if (p->dtype == D3_DTYPE_F32) {
    D3_REQUIRE(p->x && p->w && p->out && p->bias, "conv: null pointer");
    D3_REQUIRE(p->up2x_phase >= 0 && p->up2x_phase <= 4, "conv: bad up2x_phase");
    D3_REQUIRE((p->Cout_valid == 0 || p->Cout_valid == p->Cout) && !p->asym_pad,
               "conv: Cout_valid / asym_pad are 16-bit tensor-core path options");
    return conv2d_f32(p, stream);
  }
  D3_REQUIRE(p->Cin > 0 && p->Cin % 64 == 0, "conv: Cin must be a multiple of 64 (got %d)", p->Cin);
  D3_REQUIRE(p->Cout > 0 && p->Cout % 128 == 0, "conv: Cout must be a multiple of 128 (got %d)", p->Cout);
  D3_REQUIRE(p->C1 % 64 == 0 && p->C2 % 64 == 0, "conv: shortcut channels must be multiples of 64");
  D3_REQUIRE(p->x_pitch % 8 == 0 && p->out_pitch % 8 == 0, "conv: pitches must be multiples of 8");
  D3_REQUIRE(p->x && p->w && p->out && p->bias, "conv: null pointer");
  D3_REQUIRE(p->dtype == D3_DTYPE_BF16 || p->dtype == D3_DTYPE_F16, "conv: bad dtype %d", p->dtype);
  if (p->stride == 2) {
    D3_REQUIRE(p->H % 2 == 0 && p->W % 2 == 0, "conv: stride 2 needs even H, W");
    D3_REQUIRE(p->C1 == 0 && p->C2 == 0, "conv: stride 2 with shortcut not supported");
  }
  const int Ho = p->H / p->stride, Wo = p->W / p->stride;
  const int phase = p->up2x_phase;  // 0: plain conv; 1..4: sub-pixel phase (py, px) = ((phase-1)/2, (phase-1)%2)
  D3_REQUIRE(phase >= 0 && phase <= 4, "conv: bad up2x_phase %d", phase);
  D3_REQUIRE(phase == 0 || (p->ksize == 3 && p->stride == 1 && !p->sc1 && !p->sc2 && !p->residual),
             "conv: up2x_phase needs a plain 3x3 stride-1 conv");
  const int taps = phase ? 4 : p->ksize * p->ksize;
  const int Ktot = taps * p->Cin + p->C1 + p->C2;

  ConvKArgs a{};
  a.B = p->B;
  a.Ho = Ho;
  a.Wo = Wo;
  a.Cout = p->Cout;
  const int CoutV = p->Cout_valid > 0 ? p->Cout_valid : p->Cout;
  a.CoutV = CoutV;
  D3_REQUIRE(CoutV <= p->Cout && CoutV % 64 == 0, "conv: Cout_valid must be a multiple of 64, <= Cout (got %d)", CoutV);
  D3_REQUIRE(CoutV == p->Cout || (p->stats == nullptr && p->up2x_phase >= 0),
             "conv: GroupNorm statistics are not available with a clipped output (Cout_valid < Cout)");
  D3_REQUIRE(!p->asym_pad || (p->stride == 2 && p->ksize == 3), "conv: asym_pad needs a 3x3 stride-2 conv");
  // Cout == 128 layers run the pixels-as-N kernel (weights on the M side); see conv_igemm_t_kernel.
  static int no_t = -1;
  if (no_t < 0) { const char* e = getenv("D3_CONV_NO_T"); no_t = e ? atoi(e) : 0; }
  // halo-block variant: plain 3x3 stride-1 convs whose output is covered well by TH x TW = 256-pixel tiles with TW a
  // multiple of 8 (D3_CONV_VB: 0 = off, 1 = pick TW automatically, 8/16/32/64 = force that tile width)
  static int want_vb = -1;
  if (want_vb < 0) { const char* e = getenv("D3_CONV_VB"); want_vb = e ? atoi(e) : 1; }
  bool use_strips = false;
  if (!no_t && p->Cout % 128 == 0 && want_vb && p->ksize == 3 && p->stride == 1 && phase == 0) {
    double best = 0.0;
    for (int tw = (want_vb == 8 ? 8 : 16); tw <= 64; tw *= 2) {
      if (want_vb > 1 && tw != want_vb) continue;
      const int th = 256 / tw;
      const double eff = double(Ho) * Wo / (double(ceil_div(Ho, th)) * th * ceil_div(Wo, tw) * tw);
      // activation bytes per K-block relative to per-tap tiles: (th + 2) / (3 * th); prefer the smaller unless it wastes MMAs
      const double score = eff * (1.0 - 0.15 * double(th + 2) / (3.0 * th));
      if ((eff >= 0.85 || want_vb > 1) && score > best) { best = score; a.TH = th; a.TW = tw; use_strips = true; }
    }
  }
  // Wider layers run the pixels-as-N kernel too, as Cout / 128 channel tiles over the same pixel tile - but only in
  // its halo-block form and with enough 256-pixel tiles to fill the SMs: measured +5..17 % over the pixels-as-M kernel
  // on the 256-channel 3x3 layers (76 vs 96 KB of shared-memory traffic per K-block), while per-tap pixels-as-N tiles
  // are up to 2x SLOWER on the 1x1 / small / stride-2 layers (profiles/r01_conv_experiments.txt).
  // D3_CONV_T_MAXC caps the channel count (128 restores the old routing).
  static int t_maxc = -1;
  if (t_maxc < 0) { const char* e = getenv("D3_CONV_T_MAXC"); t_maxc = e ? atoi(e) : 4096; }
  const long long t_tiles = ((long long)p->B * Ho * Wo + 255) / 256 * (p->Cout / 128);
  const bool use_t = (p->Cout % 128 == 0) && !no_t &&
                     (p->Cout == 128 || (use_strips && p->Cout <= t_maxc && t_tiles >= num_sms()));
  if (!use_t) use_strips = false;
  if (use_strips) {
    a.blk_slot = (a.TH + 2) * a.TW * 128;
    // ring depths: 32 KB staging + barriers leave ~193 KB
    a.n_blk = (a.TW == 64) ? 2 : 3;
    a.n_wts = (193 * 1024 - a.n_blk * a.blk_slot) / (int)kWTileBytes;
    if (a.n_wts > kMaxWts) a.n_wts = kMaxWts;
    static int nb = -1;
    if (nb < 0) { const char* e = getenv("D3_CONV_VB_NBLK"); nb = e ? atoi(e) : 0; }
    if (nb >= 2 && nb <= kMaxBlk && (193 * 1024 - nb * a.blk_slot) / (int)kWTileBytes >= 3) {
      a.n_blk = nb;
      a.n_wts = (193 * 1024 - nb * a.blk_slot) / (int)kWTileBytes;
      if (a.n_wts > kMaxWts) a.n_wts = kMaxWts;
    }
  } else if (use_t) {
    choose_tile_t(Ho, Wo, &a.TH, &a.TW);
  } else {
    choose_tile(Ho, Wo, &a.TH, &a.TW);
  }
  a.tilesH = ceil_div(Ho, a.TH);
  a.tilesW = ceil_div(Wo, a.TW);
  const int pixTiles = p->B * a.tilesH * a.tilesW;
  int BN = (p->Cout % 256 == 0) ? 256 : 128;
  if (BN == 256 && pixTiles * (p->Cout / 256) < num_sms()) BN = 128;  // more CTAs for tiny layers
  if (use_t) BN = 128;  // pixels-as-N: the weight (M) tile is always 128 output channels
  a.nTilesN = p->Cout / BN;
  a.numTiles = pixTiles * a.nTilesN;
  a.splitK = 1;
  a.kbPer = 0;   // set below once kbTotal is known
  a.ws = nullptr;
  a.ksize = p->ksize;
  a.stride = p->stride;
  a.pad = p->asym_pad ? 0 : p->ksize / 2;
  a.ntaps = taps;
  if (phase) {
    // nearest-2x upsample folded into the conv: output pixel (2i+py, 2j+px) reads input rows {i+py-1, i+py} and
    // columns {j+px-1, j+px} with pre-summed weights (see ops.pack_upsample_weights)
    const int py = (phase - 1) / 2, px = (phase - 1) % 2;
    for (int t = 0; t < 4; ++t) {
      a.tap_dh[t] = (signed char)(py - 1 + t / 2);
      a.tap_dw[t] = (signed char)(px - 1 + t % 2);
    }
  } else {
    for (int t = 0; t < taps; ++t) {
      a.tap_dh[t] = (signed char)(t / p->ksize - a.pad);
      a.tap_dw[t] = (signed char)(t % p->ksize - a.pad);
    }
  }
  a.Cin = p->Cin;
  a.cbMain = p->Cin / 64;
  a.kbMain = taps * a.cbMain;
  a.kbSc1 = p->sc1 ? p->C1 / 64 : 0;
  a.kbSc2 = p->sc2 ? p->C2 / 64 : 0;
  a.kbTotal = a.kbMain + a.kbSc1 + a.kbSc2;
  a.kbPer = a.kbTotal;
  if (!use_t && p->workspace != nullptr) {
    const int S = plan_split_k(p, nullptr);
    if (S > 1 && (size_t)p->workspace_bytes >= (size_t)S * p->B * Ho * Wo * p->Cout * sizeof(float)) {
      a.splitK = S;
      a.kbPer = ceil_div(a.kbTotal, S);
      a.ws = static_cast<float*>(p->workspace);
    }
  }
  a.sc_off2 = p->x_pitch;
  a.bias = p->bias;
  a.bias_bstride = p->bias_bstride;
  a.residual = static_cast<const uint16_t*>(p->residual);
  a.res_pitch = p->res_pitch;
  a.stats = p->stats;
  a.a_bytes = static_cast<uint32_t>(a.TH * a.TW * 128);
  a.debug_skip = 0;
#ifdef D3_DEV
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("D3_CONV_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; }
    a.debug_skip = dbg;
  }
#endif
  D3_REQUIRE((p->sc1 ? p->C1 : 0) + (p->sc2 ? p->C2 : 0) == p->C1 + p->C2, "conv: C1/C2 set without pointers");

  CUtensorMap tmA, tmS1, tmS2, tmB, tmO, tmR;
  int rc = 0;
  a.pass_w = use_t ? pass_width(a.TW) : 0;
  a.pass_h = (use_strips && a.TW < 64) ? 64 / a.TW : 1;
  if (use_t && !use_strips)
    D3_REQUIRE(t_tile_windows_ok(a.TH, a.TW), "conv: internal error: tile %dx%d overruns the accumulator", a.TH, a.TW);
  const int store_rows = use_t ? a.pass_h : a.TH;        // rows per TMA store box
  const int store_cols = use_t ? a.pass_w : a.TW;        // pixels per row of the TMA store box
  if (p->stride == 1) {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->x_pitch * 2, (uint64_t)p->W * p->x_pitch * 2,
                       (uint64_t)p->H * p->W * p->x_pitch * 2};
    uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)(use_strips ? a.TH + 2 : a.TH), 1};
    rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmA, p->x, 4, dims, str, box);
  } else {
    // parity-split view: (pw*pitch + c, w/2, h parity, h/2, b)
    uint64_t dims[5] = {(uint64_t)p->x_pitch + p->Cin, (uint64_t)p->W / 2, 2, (uint64_t)p->H / 2, (uint64_t)p->B};
    uint64_t str[4] = {(uint64_t)p->x_pitch * 4, (uint64_t)p->W * p->x_pitch * 2,
                       (uint64_t)p->W * p->x_pitch * 4, (uint64_t)p->H * p->W * p->x_pitch * 2};
    uint32_t box[5] = {64, (uint32_t)a.TW, 1, (uint32_t)a.TH, 1};
    rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmA, p->x, 5, dims, str, box);
  }
  if (rc) return set_error(D3_ERR_CUDA, "conv: tensor map A encode failed (%d)", rc);
  tmS1 = tmA;
  tmS2 = tmA;
  if (a.kbSc1) {
    uint64_t dims[4] = {(uint64_t)p->C1, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->sc1_pitch * 2, (uint64_t)Wo * p->sc1_pitch * 2,
                       (uint64_t)Ho * Wo * p->sc1_pitch * 2};
    uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)a.TH, 1};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmS1, p->sc1, 4, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map S1 encode failed (%d)", rc);
  }
  if (a.kbSc2) {
    uint64_t dims[4] = {(uint64_t)p->C2, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->sc2_pitch * 2, (uint64_t)Wo * p->sc2_pitch * 2,
                       (uint64_t)Ho * Wo * p->sc2_pitch * 2};
    uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)a.TH, 1};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmS2, p->sc2, 4, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map S2 encode failed (%d)", rc);
  }
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)p->Cout};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmB, p->w, 2, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map B encode failed (%d)", rc);
  }
  {
    // channels >= Cout_valid are weight padding: the TMA store (and the residual load) clips them
    uint64_t dims[4] = {(uint64_t)CoutV, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->out_pitch * 2, (uint64_t)Wo * p->out_pitch * 2,
                       (uint64_t)Ho * Wo * p->out_pitch * 2};
    const void* obase = p->out;
    if (phase) {
      // strided view of the [B, 2H, 2W, Cout] output: every second pixel in both directions, offset (py, px)
      const int py = (phase - 1) / 2, px = (phase - 1) % 2;
      str[0] = (uint64_t)p->out_pitch * 4;
      str[1] = (uint64_t)(2 * Wo) * p->out_pitch * 4;
      str[2] = (uint64_t)(2 * Ho) * (2 * Wo) * p->out_pitch * 2;
      obase = static_cast<const char*>(p->out) + ((size_t)py * (2 * Wo) + px) * p->out_pitch * 2;
    }
    uint32_t box[4] = {64, (uint32_t)store_cols, (uint32_t)store_rows, 1};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmO, obase, 4, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map O encode failed (%d)", rc);
    tmR = tmO;
    if (use_t && p->residual) {
      uint64_t rstr[3] = {(uint64_t)p->res_pitch * 2, (uint64_t)Wo * p->res_pitch * 2,
                          (uint64_t)Ho * Wo * p->res_pitch * 2};
      if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmR, p->residual, 4, dims, rstr, box)))
        return set_error(D3_ERR_CUDA, "conv: tensor map R encode failed (%d)", rc);
    }
  }

  int grid = a.numTiles * a.splitK < num_sms() ? a.numTiles * a.splitK : num_sms();
  cudaError_t le = cudaSuccess;
  // epilogue of the halo-block kernel: 1 = stmatrix transposing epilogue (default), 0 = lane-exchange epilogue
  // (D3_CONV_EPI, a routing override for A/B measurements; both are parity-tested)
  static int want_epi = -1;
  if (want_epi < 0) { const char* e = getenv("D3_CONV_EPI"); want_epi = e ? atoi(e) : 1; }
  const int epi = (use_strips && want_epi) ? 1 : 0;
  if (use_t && p->dtype == D3_DTYPE_BF16)
    le = launch_conv_t<4, kBF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, tmR, a, use_strips, epi);
  else if (use_t)
    le = launch_conv_t<4, kF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, tmR, a, use_strips, epi);
  else if (BN == 128 && p->dtype == D3_DTYPE_BF16)
    le = launch_conv<128, 5, kBF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, a);
  else if (BN == 128)
    le = launch_conv<128, 5, kF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, a);
  else if (p->dtype == D3_DTYPE_BF16)
    le = launch_conv<256, 4, kBF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, a);
  else
    le = launch_conv<256, 4, kF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, a);
  D3_CHECK_CUDA(le);
  if (a.splitK > 1) {
    D3_LAUNCH_CHECK();
    SplitKFinArgs f{a.ws, a.splitK, p->bias, p->bias_bstride, static_cast<const uint16_t*>(p->residual), p->res_pitch,
                    static_cast<uint16_t*>(p->out), p->out_pitch, p->stats, p->B, Ho * Wo, p->Cout};
    const int nvec = p->Cout / 8;
    const int threads = nvec <= 256 ? nvec * (256 / nvec) : 0;
    D3_REQUIRE(threads > 0, "conv: split-K finalize supports Cout <= 2048 (got %d)", p->Cout);
    const int pstride = threads / nvec;
    int gx = ceil_div(Ho * Wo, pstride * 4);   // ~4 pixels per thread
    const int cap = ceil_div(num_sms() * 2, p->B);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 fgrid(gx, p->B);
    if (p->dtype == D3_DTYPE_BF16)
      D3_CHECK_CUDA(launch_pdl(splitk_finalize_kernel<kBF16>, fgrid, dim3(threads), 0, stream, f));
    else
      D3_CHECK_CUDA(launch_pdl(splitk_finalize_kernel<kF16>, fgrid, dim3(threads), 0, stream, f));
  }
  D3_LAUNCH_CHECK();
  return D3_OK;
}
