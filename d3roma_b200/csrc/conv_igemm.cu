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

namespace d3 {

namespace {

constexpr int kThreads = 256;
constexpr int kTileM = 128;
constexpr uint32_t kATileBytes = kTileM * 64 * 2;  // 16 KB
constexpr uint32_t kStagingBytes = kTileM * 128;   // one 64-channel chunk of the output tile

struct ConvKArgs {
  int B, Ho, Wo, Cout;
  int TH, TW, tilesH, tilesW, nTilesN, numTiles;
  int numPairs;  // cta_group::2 kernel: pairs of pixel tiles x nTilesN
  int numUnits, pixTiles;  // weight-multicast clusters: work units = pairs of pixel tiles x nTilesN
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
  int t_direct;    // pixels-as-N kernel: epilogue stores straight to global (no staging / TMA store)
  uint16_t* out;   // output base pointer + pitch for the direct epilogue
  int out_pitch;
  int pass_w;      // pixels-as-N kernel: pixels per epilogue pass (divides TW, <= 64)
  int debug_skip;  // developer experiment: bit0 = skip A loads, bit1 = skip B loads (results are garbage)
};

struct TileCoord {
  int b, h0, w0, n0;
};

// developer profiling (D3_CONV_DEBUG_SKIP bit 256): clock cycles summed over CTAs
//   [0] MMA thread total, [1] waiting for a free accumulator, [2] waiting for pixel blocks, [3] waiting for weights,
//   [4] epilogue thread total, [5] waiting for a full accumulator, [6] inside passes, [7] tiles
//   [8] pass: store-buffer wait + barrier 1, [9] TMEM load, [10] math + staging stores, [11] fence + barrier 2 + store issue
__device__ unsigned long long g_conv_prof[16];

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

// Work decode of the pixels-as-N kernel.  MC (2-CTA weight-multicast clusters): unit = (pair of pixel tiles, channel
// tile); CTA `rank` of the cluster takes pixel tile 2*pair + rank; an odd tile count leaves one CTA with a dummy tile
// (valid = false: it still consumes the multicast weights, loads zeros and stores nothing).
template <bool MC>
D3_DEVINL TileCoord decode_work(const ConvKArgs& a, int idx, uint32_t rank, bool& valid) {
  if constexpr (!MC) {
    valid = true;
    return decode_tile(a, idx, 128);
  } else {
    TileCoord t;
    const int nt = idx % a.nTilesN;
    int p = 2 * (idx / a.nTilesN) + (int)rank;
    valid = p < a.pixTiles;
    t.n0 = nt * 128;
    if (!valid) {
      t.b = a.B;  // out of range in the batch dimension: TMA loads return zeros, TMA stores are dropped
      t.h0 = 0;
      t.w0 = 0;
      return t;
    }
    const int tw = p % a.tilesW;
    p /= a.tilesW;
    t.h0 = (p % a.tilesH) * a.TH;
    t.b = p / a.tilesH;
    t.w0 = tw * a.TW;
    return t;
  }
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
          if (a.debug_skip) {
            const uint32_t bytes = ((a.debug_skip & 1) ? 0u : a.a_bytes) + ((a.debug_skip & 2) ? 0u : kBBytes);
            if (bytes) mbar_arrive_expect_tx(&full_bar[stage], bytes); else mbar_arrive(&full_bar[stage]);
            if (!(a.debug_skip & 1)) tma_load_4d(sA, &tmA, &full_bar[stage], 0, t.w0, t.h0, t.b);
            if (!(a.debug_skip & 2)) tma_load_2d(sB, &tmB, &full_bar[stage], kb * 64, t.n0);
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
              // input pixel (2*ho + kh - 1, 2*wo + kw - 1) in the parity-split 5-D view
              const int dh = kh - 1, dw = kw - 1;
              const int hh = t.h0 + (dh < 0 ? -1 : 0), ph = (dh == 0) ? 0 : 1;
              const int ww = t.w0 + (dw < 0 ? -1 : 0), pw = (dw == 0) ? 0 : 1;
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
      if (a.debug_skip & 4) {  // developer experiment: no epilogue work at all
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
        if (a.residual != nullptr && valid) {
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
          tma_store_4d(&tmO, sbuf, n_base, t.w0, t.h0, t.b);
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


// One 32-pixel chunk of the pixels-as-N epilogue for one channel (thread): + bias (+ residual read back from the
// staging tile), round to 16 bit, store at [pixel row][channel] of the SWIZZLE_128B staging tile, accumulate the
// channel's sum / sum-of-squares over in-image pixels.
template <int DT, bool HAS_RES, int J = 0>
D3_DEVINL void epi_t_chunk(const uint32_t (&v)[32], uint32_t sa, const uint32_t (&swz)[8], float bias, uint32_t vmask,
                           float& s_sum, float& s_sq) {
  if constexpr (J < 32) {
    const uint32_t addr = sa + swz[J & 7];
    float x = __uint_as_float(v[J]) + bias;
    if constexpr (HAS_RES) x += unpack2<DT>(lds_u16<J * 128>(addr)).x;
    const uint32_t pk = pack2<DT>(x, 0.f);
    sts_u16<J * 128>(addr, pk);
    const float xr = ((vmask >> J) & 1u) ? unpack2<DT>(pk & 0xFFFFu).x : 0.f;
    s_sum += xr;
    s_sq = fmaf(xr, xr, s_sq);
    epi_t_chunk<DT, HAS_RES, J + 1>(v, sa, swz, bias, vmask, s_sum, s_sq);
  }
}

// Paired variant of the chunk epilogue: lanes 2k / 2k+1 (adjacent channels) exchange values so that the even lane
// stores the 32-bit word {ch, ch+1} of pixel row j and the odd lane the word {ch-1, ch} of pixel row j+4.  Rows j and
// j+4 sit in opposite halves of the swizzle pattern, so a warp's 32 word stores hit 32 distinct banks (the 16-bit
// version measured 50 % bank conflicts).  Statistics: each lane sums both halves of the words it stores; the quad
// total is the sum over the 4 lanes of a quad, which is all the caller needs.
template <int DT, bool HAS_RES, int BASE = 0, int JJ = 0, bool NO_STORE = false>
D3_DEVINL void epi_t_chunk2(const uint32_t (&v)[32], uint32_t sa_lane, const uint32_t (&myswz)[4], float bias,
                            uint32_t vmask, bool odd, float& s_sum, float& s_sq) {
  if constexpr (BASE < 32) {
    constexpr int J0 = BASE + JJ, J1 = J0 + 4;
    const float x0 = __uint_as_float(v[J0]) + bias;   // own channel, pixel J0
    const float x1 = __uint_as_float(v[J1]) + bias;   // own channel, pixel J1
    const float recv = __shfl_xor_sync(0xffffffffu, odd ? x0 : x1, 1);
    float lo = odd ? recv : x0;                       // even: {ch @J0, ch+1 @J0}; odd: {ch-1 @J1, ch @J1}
    float hi = odd ? x1 : recv;
    const uint32_t addr = sa_lane + myswz[JJ];
    if constexpr (HAS_RES) {
      const float2 r = unpack2<DT>(lds_u32<J0 * 128>(addr));
      lo += r.x;
      hi += r.y;
    }
    const uint32_t pk = pack2<DT>(lo, hi);
    if (!NO_STORE) sts_u32<J0 * 128>(addr, pk);
    const bool ok = ((vmask >> (odd ? J1 : J0)) & 1u) != 0;
    const float2 xr = unpack2<DT>(pk);
    const float a0 = ok ? xr.x : 0.f, a1 = ok ? xr.y : 0.f;
    s_sum += a0 + a1;
    s_sq = fmaf(a0, a0, fmaf(a1, a1, s_sq));
    if constexpr (JJ < 3)
      epi_t_chunk2<DT, HAS_RES, BASE, JJ + 1, NO_STORE>(v, sa_lane, myswz, bias, vmask, odd, s_sum, s_sq);
    else
      epi_t_chunk2<DT, HAS_RES, BASE + 8, 0, NO_STORE>(v, sa_lane, myswz, bias, vmask, odd, s_sum, s_sq);
  }
}

// Packed fp32x2 accumulate (Blackwell FADD2 / FFMA2): acc += x, sq += x * x on two lanes of a register pair.
D3_DEVINL void stats_acc2(float2& s, float2& q, float2 x) {
  asm("{\n\t.reg .b64 s, q, x;\n\t"
      "mov.b64 s, {%0, %1};\n\tmov.b64 q, {%2, %3};\n\tmov.b64 x, {%4, %5};\n\t"
      "add.rn.f32x2 s, s, x;\n\tfma.rn.f32x2 q, x, x, q;\n\t"
      "mov.b64 {%0, %1}, s;\n\tmov.b64 {%2, %3}, q;\n\t}"
      : "+f"(s.x), "+f"(s.y), "+f"(q.x), "+f"(q.y)
      : "f"(x.x), "f"(x.y));
}

// Lean chunk epilogue (default): each lane rounds ITS OWN channel at pixels J0 and J0+4 into one packed word, swaps it
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

// Variant that writes straight to global memory (no staging tile, no TMA store): the warp's 32 lanes hold 32
// consecutive channels of one pixel, i.e. each store instruction covers one contiguous 64-byte segment.
template <int DT, bool HAS_RES, int J = 0>
D3_DEVINL void epi_t_chunk_global(const uint32_t (&v)[32], uint16_t* gp, const uint16_t* rp, size_t pitch, float bias,
                                  uint32_t vmask, float& s_sum, float& s_sq) {
  if constexpr (J < 32) {
    if ((vmask >> J) & 1u) {
      float x = __uint_as_float(v[J]) + bias;
      if constexpr (HAS_RES) x += unpack2<DT>((uint32_t)__ldg(rp + (size_t)J * pitch)).x;
      const uint32_t pk = pack2<DT>(x, 0.f);
      gp[(size_t)J * pitch] = (uint16_t)pk;
      const float xr = unpack2<DT>(pk & 0xFFFFu).x;
      s_sum += xr;
      s_sq = fmaf(xr, xr, s_sq);
    }
    epi_t_chunk_global<DT, HAS_RES, J + 1>(v, gp, rp, pitch, bias, vmask, s_sum, s_sq);
  }
}

constexpr int kTThreads = 384;  // warps 0-3: producer / MMA / TMEM alloc / idle; warps 4-11: two epilogue groups

constexpr int kMaxBlk = 3;    // halo-block ring slots (halo-block variant)
constexpr int kMaxWts = 8;    // weight-tile ring slots (halo-block variant)

// STRIPS = true ("halo blocks"): 3x3 stride-1 convs.  The pixel operand of a TH x TW tile (TW a multiple of 8, so one
// tile row = TW*128 B is a whole number of 1024 B swizzle atoms) is loaded ONCE per (kw, 64-channel block) as a block of
// TH+2 rows starting one row above the tile, and the three kh taps read it through UMMA descriptors whose start
// address is advanced by whole tile rows - atom-aligned, unlike a kw (single-pixel) shift, which measured ~25 % slower
// in the MMA itself (profiles/r01_conv_experiments.txt).  Activation TMA traffic drops by 3*TH/(TH+2).  Weights stream
// through their own ring, fed by a second producer lane (warp 3).
// K-loop order: for cb: for kw: block(kw, cb) -> for kh: weight tile (kh, kw, cb) -> 4 MMAs.
template <int STAGES, int DT, bool STRIPS, bool MC = false>
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
      // MC: a weight slot is free again only when BOTH CTAs of the cluster have consumed it (each producer multicasts
      // its half of the tile into both CTAs)
      mbar_init(&empty_bar[i], (MC && STRIPS && i >= kMaxBlk) ? 2 : 1);
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
  if constexpr (MC) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  const int w_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // work index of this CTA / cluster
  const int w_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_count = MC ? a.numUnits : a.numTiles;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (STRIPS && lane == 0) {
      int ss = 0;
      uint32_t sphase = 0;
      const uint32_t blk_bytes = (uint32_t)(a.TH + 2) * a.TW * 128u;
      const uint32_t plain_bytes = (uint32_t)a.TH * a.TW * 128u;
      const int n_items = 3 * a.cbMain + a.kbSc1 + a.kbSc2;  // block loads per tile
      for (int tile = w_first; tile < w_count; tile += w_step) {
        bool valid;
        const TileCoord t = decode_work<MC>(a, tile, crank, valid);
        for (int it = 0; it < n_items; ++it) {
          mbar_wait(&sempty[ss], sphase ^ 1);
          uint8_t* sS = smem + ss * a.blk_slot;
          if (a.debug_skip & 1) {  // developer experiment: no activation loads
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
          if (a.debug_skip & 3) {
            const uint32_t bytes = ((a.debug_skip & 1) ? 0u : a.a_bytes) + ((a.debug_skip & 2) ? 0u : kWTileBytes);
            if (bytes) mbar_arrive_expect_tx(&full_bar[stage], bytes); else mbar_arrive(&full_bar[stage]);
            if (!(a.debug_skip & 1)) tma_load_4d(sP, &tmA, &full_bar[stage], 0, t.w0, t.h0, t.b);
            if (!(a.debug_skip & 2)) tma_load_2d(sW, &tmB, &full_bar[stage], kb * 64, t.n0);
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
              const int dh = kh - 1, dw = kw - 1;
              const int hh = t.h0 + (dh < 0 ? -1 : 0), ph = (dh == 0) ? 0 : 1;
              const int ww = t.w0 + (dw < 0 ? -1 : 0), pw = (dw == 0) ? 0 : 1;
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
      const bool prof = (a.debug_skip & 256) != 0;
      long long p_t0 = prof ? clock64() : 0, p_acc = 0, p_blk = 0, p_wts = 0, p_c = 0;
      for (int tile = w_first; tile < w_count; tile += w_step) {
        if (prof) p_c = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (prof) p_acc += clock64() - p_c;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int it = 0; it < n_items; ++it) {
          if (prof) p_c = clock64();
          mbar_wait(&sfull[ss], sphase);
          if (prof) p_blk += clock64() - p_c;
          tc_fence_after();
          const uint32_t sS = smem_u32(smem + ss * a.blk_slot);
          const int nkb = (it < 3 * a.cbMain) ? 3 : 1;
          for (int i = 0; i < nkb; ++i) {
            if (prof) p_c = clock64();
            mbar_wait(&wfull[ws], wphase);
            if (prof) p_wts += clock64() - p_c;
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(wring + ws * kWTileBytes));
            // pixel window of tap kh = i: tile rows i .. i+TH-1 of the block (whole swizzle atoms further down)
            const uint64_t bdesc = make_sw128_kmajor_desc(sS + (uint32_t)i * row_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
            }
            first = 0;
            if constexpr (MC) umma_commit_mc(&wempty[ws], 3);   // frees the slot in both CTAs' weight producers
            else umma_commit(&wempty[ws]);
            if (++ws == a.n_wts) { ws = 0; wphase ^= 1; }
          }
          umma_commit(&sempty[ss]);
          if (++ss == a.n_blk) { ss = 0; sphase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (prof) {
        atomicAdd(&g_conv_prof[0], (unsigned long long)(clock64() - p_t0));
        atomicAdd(&g_conv_prof[1], (unsigned long long)p_acc);
        atomicAdd(&g_conv_prof[2], (unsigned long long)p_blk);
        atomicAdd(&g_conv_prof[3], (unsigned long long)p_wts);
      }
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
      for (int tile = w_first; tile < w_count; tile += w_step) {
        bool valid;
        const TileCoord t = decode_work<MC>(a, tile, crank, valid);
        for (int it = 0; it < n_items; ++it) {
          int nkb = 1, kb0 = a.kbMain + (it - 3 * a.cbMain);
          if (it < 3 * a.cbMain) {
            const int cb = it / 3, kw = it - cb * 3;
            nkb = 3;
            kb0 = kw * a.cbMain + cb;  // tap (kh, kw) -> K-block (kh*3 + kw) * cbMain + cb
          }
          for (int i = 0; i < nkb; ++i) {
            mbar_wait(&wempty[ws], wphase ^ 1);
            if (a.debug_skip & 2) {
              mbar_arrive(&wfull[ws]);
            } else {
              mbar_arrive_expect_tx(&wfull[ws], kWTileBytes);   // own half + the peer's half
              if constexpr (MC)
                tma_load_2d_mc(wring + ws * kWTileBytes + crank * (kWTileBytes / 2), &tmB, &wfull[ws],
                               (kb0 + i * 3 * a.cbMain) * 64, t.n0 + (int)crank * 64, 3);   // tmB box = 64 rows
              else
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
    const bool has_res = a.residual != nullptr;
    // A pass = PW (<= 64) consecutive pixels of one tile row = a contiguous range of accumulator columns.  Staging is
    // double-buffered by pass so the TMA store of pass p overlaps the math of pass p+1.
    const int PW = a.pass_w;
    const int PH = a.pass_h;                    // > 1 only for narrow tiles (TW < 64): a pass = PH whole tile rows
    const int PN = PW * PH;                     // pixels (accumulator columns) per pass
    const int segs = a.TW / PW;
    const int passes = (a.TH / PH) * segs;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    uint32_t pass_ctr = 0;
    const bool prof = (a.debug_skip & 256) != 0 && et == 0;
    long long p_t0 = prof ? clock64() : 0, p_wait = 0, p_pass = 0, p_c = 0, p_tiles = 0;
    long long p_q[4] = {0, 0, 0, 0};
    for (int tile = w_first; tile < w_count; tile += w_step) {
      bool valid;
      const TileCoord t = decode_work<MC>(a, tile, crank, valid);
      const float bias = valid ? a.bias[(size_t)t.b * a.bias_bstride + t.n0 + ch] : 0.f;
      float s_sum = 0.f, s_sq = 0.f;
      if (prof) p_c = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (prof) { const long long n = clock64(); p_wait += n - p_c; p_c = n; ++p_tiles; }
      tc_fence_after();
      if (a.debug_skip & 4) {  // developer experiment: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (a.t_direct) {
        // direct-to-global epilogue: 32-pixel chunks of the tile's accumulator columns, alternating between groups
        const int NPx = a.TH * a.TW;
        for (int c0 = (int)grp * 32; c0 < NPx; c0 += 64) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * 256 + c0, v);
          const int px = c0 + (int)lane;
          const int pr = px / a.TW;
          const int pw = px - pr * a.TW;
          const bool ok = (px < NPx) && (t.h0 + pr < a.Ho) && (t.w0 + pw < a.Wo);
          const uint32_t vmask = __ballot_sync(0xffffffffu, ok);
          // pixel offset (in pixels) of chunk element j relative to element 0: contiguous inside a tile row
          const int r0 = c0 / a.TW, w0c = c0 - r0 * a.TW;
          const size_t pix0 = ((size_t)(t.b * a.Ho + t.h0 + r0) * a.Wo + t.w0 + w0c);
          tmem_ld_wait();
          if (w0c + 32 <= a.TW) {  // the 32 pixels of this chunk lie in one tile row: constant pixel stride
            uint16_t* gp = a.out + pix0 * a.out_pitch + t.n0 + ch;
            if (has_res) {
              const uint16_t* rp = a.residual + pix0 * a.res_pitch + t.n0 + ch;
              if (a.res_pitch == a.out_pitch)
                epi_t_chunk_global<DT, true>(v, gp, rp, (size_t)a.out_pitch, bias, vmask, s_sum, s_sq);
              else
                for (int j = 0; j < 32; ++j)
                  if ((vmask >> j) & 1u) {
                    const float x = __uint_as_float(v[j]) + bias + unpack2<DT>((uint32_t)rp[(size_t)j * a.res_pitch]).x;
                    const uint32_t pk = pack2<DT>(x, 0.f);
                    gp[(size_t)j * a.out_pitch] = (uint16_t)pk;
                    const float xr = unpack2<DT>(pk & 0xFFFFu).x;
                    s_sum += xr; s_sq = fmaf(xr, xr, s_sq);
                  }
            } else {
              epi_t_chunk_global<DT, false>(v, gp, nullptr, (size_t)a.out_pitch, bias, vmask, s_sum, s_sq);
            }
          } else {  // chunk straddles tile rows (narrow tiles): generic addressing
            for (int j = 0; j < 32; ++j) {
              if ((vmask >> j) & 1u) {
                const int pxj = c0 + j, prj = pxj / a.TW, pwj = pxj - prj * a.TW;
                const size_t pix = ((size_t)(t.b * a.Ho + t.h0 + prj) * a.Wo + t.w0 + pwj);
                float x = __uint_as_float(v[j]) + bias;
                if (has_res) x += unpack2<DT>((uint32_t)a.residual[pix * a.res_pitch + t.n0 + ch]).x;
                const uint32_t pk = pack2<DT>(x, 0.f);
                a.out[pix * a.out_pitch + t.n0 + ch] = (uint16_t)pk;
                const float xr = unpack2<DT>(pk & 0xFFFFu).x;
                s_sum += xr; s_sq = fmaf(xr, xr, s_sq);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      } else
      for (int ps = 0; ps < passes; ++ps, ++pass_ctr) {
        const int r = (ps / segs) * PH;
        const int sg = ps - (ps / segs) * segs;
        const int h = t.h0 + r;
        const int wpass = t.w0 + sg * PW;
        uint8_t* sbuf = staging + (pass_ctr & 1) * kStagingBytes;  // two 64-channel halves of 64 rows each
        long long q0 = prof ? clock64() : 0, q1 = 0, q2 = 0, q3 = 0;
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
        if (prof) q1 = clock64();
        if (c0 < PN) {
          uint32_t v[32];
          if (!(a.debug_skip & 64))
            tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * 256 + r * a.TW + sg * PW + c0, v);
          const int pc = c0 + (int)lane;          // pixel of the pass; row pc / PW, column pc % PW (PH == 1: row 0)
          const int prow = (PH > 1) ? pc / PW : 0;
          const int pcol = pc - prow * PW;
          const uint32_t vmask = __ballot_sync(
              0xffffffffu, valid && (pc < PN) && (h + prow < a.Ho) && (wpass + pcol < a.Wo));
          // word address of this lane's channel pair; odd lanes write 4 pixel rows further down (see epi_t_chunk2)
          const uint32_t sa = smem_u32(sbuf) + half * (kStagingBytes / 2) + (uint32_t)c0 * 128u + (chb & ~3u) +
                              (odd ? 512u : 0u);
          // byte address of this lane's own channel at pixel row 0 of the chunk (residual read-back)
          const uint32_t sa_own = smem_u32(sbuf) + half * (kStagingBytes / 2) + (uint32_t)c0 * 128u + chb;
          tmem_ld_wait();
          if (prof) q2 = clock64();
          if (a.debug_skip & 32) {  // developer experiment: no epilogue math
          } else if (a.debug_skip & 8) {
            epi_t_chunk2<DT, false, 0, 0, true>(v, sa, myswz, bias, vmask, odd, s_sum, s_sq);
          } else if (a.debug_skip & 16) {  // previous (select-based) epilogue, kept for A/B timing
            if (has_res) epi_t_chunk2<DT, true>(v, sa, myswz, bias, vmask, odd, s_sum, s_sq);
            else epi_t_chunk2<DT, false>(v, sa, myswz, bias, vmask, odd, s_sum, s_sq);
          } else {
            float2 s2 = make_float2(s_sum, 0.f), q2 = make_float2(s_sq, 0.f);
            const bool full = (vmask == 0xffffffffu);
            if (a.debug_skip & 128) {  // developer experiment: no lane exchange (wrong results)
              epi_t_chunk3<DT, false, false, 0, 0, true>(v, sa, myswz, sa_own, swz, bias, vmask, prmt_sel, s2, q2);
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
        if (prof) q3 = clock64();
        if (ps == passes - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (et == 0 && !(a.debug_skip & 8)) {
          tma_store_4d(&tmO, sbuf, t.n0, wpass, h, t.b);
          tma_store_4d(&tmO, sbuf + kStagingBytes / 2, t.n0 + 64, wpass, h, t.b);
          tma_store_commit();
        }
        if (prof) {
          const long long q4 = clock64();
          p_q[0] += q1 - q0; p_q[1] += q2 - q1; p_q[2] += q3 - q2; p_q[3] += q4 - q3;
        }
      }
      if (prof) p_pass += clock64() - p_c;
      if (a.stats != nullptr && valid) {
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
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
    if (prof) {
      atomicAdd(&g_conv_prof[4], (unsigned long long)(clock64() - p_t0));
      atomicAdd(&g_conv_prof[5], (unsigned long long)p_wait);
      atomicAdd(&g_conv_prof[6], (unsigned long long)p_pass);
      atomicAdd(&g_conv_prof[7], (unsigned long long)p_tiles);
      for (int i = 0; i < 4; ++i) atomicAdd(&g_conv_prof[8 + i], (unsigned long long)p_q[i]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();   // no CTA leaves while its peer can still multicast into it / arrive on its barriers
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
                          const CUtensorMap& tmR, const ConvKArgs& a, bool strips, bool mc) {
  if (strips && mc) {
    // 2-CTA clusters sharing every weight tile through TMA multicast; grid = 2 x clusters
    const size_t smem = conv_ts_smem_bytes(a);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(conv_igemm_t_kernel<STAGES, DT, true, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_smem = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv_igemm_t_kernel<STAGES, DT, true, true>, tmA, tmS1, tmS2, tmB, tmO, tmR, a);
  }
  if (strips) {
    const size_t smem = conv_ts_smem_bytes(a);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(conv_igemm_t_kernel<STAGES, DT, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_smem = smem;
    }
    conv_igemm_t_kernel<STAGES, DT, true><<<grid, kTThreads, smem, stream>>>(tmA, tmS1, tmS2, tmB, tmO, tmR, a);
    return cudaSuccess;
  }
  constexpr size_t smem = conv_t_smem_bytes<STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_t_kernel<STAGES, DT, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  conv_igemm_t_kernel<STAGES, DT, false><<<grid, kTThreads, smem, stream>>>(tmA, tmS1, tmS2, tmB, tmO, tmR, a);
  return cudaSuccess;
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


// =====================================================================================================================
// CTA-pair variant (cta_group::2) of the pixels-as-M kernel: two CTAs of a cluster own two adjacent pixel tiles
// (M = 256 = 2 x 128 rows) and each keeps only HALF of the weight tile (BN/2 output channels) in shared memory; one
// tcgen05.mma.cta_group::2 issued by the leader drives both SMs.  Per CTA the weight operand costs half the TMA
// writes and half the operand reads of the single-CTA kernel, which is what the shared-memory port is short of
// (profiles/r01_conv_experiments.txt).
// =====================================================================================================================
template <int BN, int STAGES, int DT>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmS1,
                   const __grid_constant__ CUtensorMap tmS2, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const ConvKArgs a) {
  constexpr uint32_t kBHalfBytes = (BN / 2) * 128;
  constexpr uint32_t kStageBytes = kATileBytes + kBHalfBytes;
  constexpr uint32_t kTmemCols = 2 * BN;

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
  const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int npairs_grid = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (a.kbSc1) tma_prefetch_desc(&tmS1);
    if (a.kbSc2) tma_prefetch_desc(&tmS2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's producer arrival (+ the bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);  // multicast commit from the leader's MMA lane
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);   // multicast commit
      mbar_init(&tempty_bar[i], 8);  // 4 epilogue warps of each CTA (only the leader's copy is waited on)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // pair-tile decode: pixel tiles 2*pp + rank share the weight tile nt
  auto decode = [&](int tp, TileCoord& t) -> bool {
    const int nt = tp % a.nTilesN;
    const int pp = tp / a.nTilesN;
    int p = 2 * pp + (int)rank;
    const int pixTiles = a.B * a.tilesH * a.tilesW;
    const bool real = p < pixTiles;
    if (!real) p = pixTiles;  // decodes to b == B: every TMA coordinate out of range (zero fill / clipped store)
    const int tw = p % a.tilesW;
    const int q = p / a.tilesW;
    const int th = q % a.tilesH;
    t.b = q / a.tilesH;
    t.h0 = th * a.TH;
    t.w0 = tw * a.TW;
    t.n0 = nt * BN;
    return real;
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tp = pair; tp < a.numPairs; tp += npairs_grid) {
        TileCoord t;
        decode(tp, t);
        for (int kb = 0; kb < a.kbTotal; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kATileBytes;
          if (a.debug_skip & 3) {
            const uint32_t bytes = ((a.debug_skip & 1) ? 0u : a.a_bytes) + ((a.debug_skip & 2) ? 0u : kBHalfBytes);
            if (rank == 0) { if (bytes) mbar_arrive_expect_tx(&full_bar[stage], 2u * bytes); else mbar_arrive(&full_bar[stage]); }
            if (!(a.debug_skip & 1)) tma_load_4d_2cta(sA, &tmA, &full_bar[stage], 0, t.w0, t.h0, t.b);
            if (!(a.debug_skip & 2)) tma_load_2d_2cta(sB, &tmB, &full_bar[stage], kb * 64, t.n0 + (int)rank * (BN / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (a.a_bytes + kBHalfBytes));
          if (kb < a.kbMain) {
            const int tap = kb / a.cbMain;
            const int cb = kb - tap * a.cbMain;
            if (a.stride == 1) {
              tma_load_4d_2cta(sA, &tmA, &full_bar[stage], cb * 64, t.w0 + a.tap_dw[tap], t.h0 + a.tap_dh[tap], t.b);
            } else {
              const int kh = tap / a.ksize;
              const int kw = tap - kh * a.ksize;
              const int dh = kh - 1, dw = kw - 1;
              const int hh = t.h0 + (dh < 0 ? -1 : 0), ph = (dh == 0) ? 0 : 1;
              const int ww = t.w0 + (dw < 0 ? -1 : 0), pw = (dw == 0) ? 0 : 1;
              tma_load_5d_2cta(sA, &tmA, &full_bar[stage], pw * a.sc_off2 + cb * 64, ww, ph, hh, t.b);
            }
          } else if (kb < a.kbMain + a.kbSc1) {
            tma_load_4d_2cta(sA, &tmS1, &full_bar[stage], (kb - a.kbMain) * 64, t.w0, t.h0, t.b);
          } else {
            tma_load_4d_2cta(sA, &tmS2, &full_bar[stage], (kb - a.kbMain - a.kbSc1) * 64, t.w0, t.h0, t.b);
          }
          // this CTA's half of the weight tile: output channels n0 + rank*BN/2 ...
          tma_load_2d_2cta(sB, &tmB, &full_bar[stage], kb * 64, t.n0 + (int)rank * (BN / 2));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_16bit(256, BN, DT == kBF16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tp = pair; tp < a.numPairs; tp += npairs_grid) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < a.kbTotal; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sA);
          const uint64_t bdesc = make_sw128_kmajor_desc(sA + kATileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta_mc(&empty_bar[stage], 0x3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta_mc(&tfull_bar[acc], 0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs, own 128 rows) ================================
    const uint32_t q = warp - 4;
    const uint32_t et = threadIdx.x - 128;
    const uint32_t r = q * 32 + lane;
    const int th = r / a.TW;
    const int tw = r - th * a.TW;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_leader1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    for (int tp = pair; tp < a.numPairs; tp += npairs_grid) {
      TileCoord t;
      const bool real = decode(tp, t);
      const int h = t.h0 + th, w = t.w0 + tw;
      const bool valid = real && (th < a.TH) && (h < a.Ho) && (w < a.Wo);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (a.debug_skip & 4) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
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
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
        }
        uint8_t* sbuf = staging + (chunk_ctr & 1) * kStagingBytes;
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        const int n_base = t.n0 + c * 64;
        const int bb = real ? t.b : 0;
        const float4* bp = reinterpret_cast<const float4*>(a.bias + (size_t)bb * a.bias_bstride + n_base);
        const uint4* rp = nullptr;
        if (a.residual != nullptr && valid) {
          rp = reinterpret_cast<const uint4*>(a.residual + ((size_t)(t.b * a.Ho + h) * a.Wo + w) * a.res_pitch +
                                              n_base);
        }
        float sums[32];
        epi_process32<0, DT>(v0, sums, bp, rp, valid, sbuf, r);
        epi_process32<1, DT>(v1, sums, bp, rp, valid, sbuf, r);
        if (a.stats != nullptr) {
          const float tot = warp_transpose_reduce32(sums, lane);
          spart[q * 32 + lane] = tot;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (et == 0) {
          tma_store_4d(&tmO, sbuf, n_base, t.w0, t.h0, t.b);
          tma_store_commit();
        }
        if (a.stats != nullptr && et < 32 && real) {
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
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES>
constexpr size_t conv2_smem_bytes() {
  return 1024 + STAGES * (kATileBytes + (BN / 2) * 128) + 2 * kStagingBytes + (2 * STAGES + 4) * 8 + 16 + 4 * 32 * 4;
}

template <int BN, int STAGES, int DT>
cudaError_t launch_conv2(int npairs_grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmS1,
                         const CUtensorMap& tmS2, const CUtensorMap& tmB, const CUtensorMap& tmO, const ConvKArgs& a) {
  constexpr size_t smem = conv2_smem_bytes<BN, STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm2_kernel<BN, STAGES, DT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * npairs_grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<BN, STAGES, DT>, tmA, tmS1, tmS2, tmB, tmO, a);
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES, DT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  conv_igemm_kernel<BN, STAGES, DT><<<grid, kThreads, smem, stream>>>(tmA, tmS1, tmS2, tmB, tmO, a);
  return cudaSuccess;
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

// Developer hook (not part of include/d3roma_b200.h): read (and optionally clear) the cycle counters collected while
// D3_CONV_DEBUG_SKIP has bit 256 set.
extern "C" int d3_debug_conv_profile(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, d3::g_conv_prof, sizeof(d3::g_conv_prof)) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(d3::g_conv_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}

extern "C" int d3_conv2d_igemm(const d3_conv_params* p, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(p != nullptr, "conv: null params");
  D3_REQUIRE(p->ksize == 3 || p->ksize == 1, "conv: ksize must be 1 or 3 (got %d)", p->ksize);
  D3_REQUIRE(p->stride == 1 || (p->stride == 2 && p->ksize == 3), "conv: unsupported stride %d", p->stride);
  if (p->dtype == D3_DTYPE_F32) {
    D3_REQUIRE(p->x && p->w && p->out && p->bias, "conv: null pointer");
    D3_REQUIRE(p->up2x_phase >= 0 && p->up2x_phase <= 4, "conv: bad up2x_phase");
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
  a.pad = p->ksize / 2;
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
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("D3_CONV_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; }
    a.debug_skip = dbg;
  }
  D3_REQUIRE((p->sc1 ? p->C1 : 0) + (p->sc2 ? p->C2 : 0) == p->C1 + p->C2, "conv: C1/C2 set without pointers");

  CUtensorMap tmA, tmS1, tmS2, tmB, tmO, tmR;
  int rc = 0;
  a.pass_w = use_t ? pass_width(a.TW) : 0;
  a.pass_h = (use_strips && a.TW < 64) ? 64 / a.TW : 1;
  if (use_t && !use_strips)
    D3_REQUIRE(t_tile_windows_ok(a.TH, a.TW), "conv: internal error: tile %dx%d overruns the accumulator", a.TH, a.TW);
  {
    static int td = -1;
    if (td < 0) { const char* e = getenv("D3_CONV_T_DIRECT"); td = e ? atoi(e) : 0; }
    a.t_direct = (use_t && phase == 0) ? td : 0;
  }
  a.out = static_cast<uint16_t*>(p->out);
  a.out_pitch = p->out_pitch;
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
    uint64_t dims[4] = {(uint64_t)p->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p->B};
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
  // weight multicast: 2-CTA clusters of the halo-block kernel share every weight tile (each CTA fetches half of it and
  // multicasts) - halves the L2 -> SM weight traffic.  D3_CONV_MC=1 to enable.
  static int want_mc = -1;
  if (want_mc < 0) { const char* e = getenv("D3_CONV_MC"); want_mc = e ? atoi(e) : 0; }
  const bool use_mc = use_strips && want_mc && !a.t_direct && pixTiles >= 4;
  if (use_mc) {
    a.pixTiles = pixTiles;
    a.numUnits = ((pixTiles + 1) / 2) * a.nTilesN;
    const int max_clusters = num_sms() / 2;
    grid = 2 * (a.numUnits < max_clusters ? a.numUnits : max_clusters);
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)p->Cout};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, 64};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmB, p->w, 2, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map B(multicast half) encode failed (%d)", rc);
  }
  static int use_2cta = -1;
  if (use_2cta < 0) { const char* e = getenv("D3_CONV_2CTA"); use_2cta = e ? atoi(e) : 0; }
  a.numPairs = ((pixTiles + 1) / 2) * a.nTilesN;
  if (!use_t && use_2cta && a.numPairs >= 8 && a.splitK == 1) {
    // B-operand box covers this CTA's half of the weight tile
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)p->Cout};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)(BN / 2)};
    if ((rc = make_tmap_16bit(p->dtype == D3_DTYPE_F16, &tmB, p->w, 2, dims, str, box)))
      return set_error(D3_ERR_CUDA, "conv: tensor map B(half) encode failed (%d)", rc);
    const int max_pairs = num_sms() / 2;
    const int npg = a.numPairs < max_pairs ? a.numPairs : max_pairs;
    if (BN == 128 && p->dtype == D3_DTYPE_BF16)
      le = launch_conv2<128, 6, kBF16>(npg, stream, tmA, tmS1, tmS2, tmB, tmO, a);
    else if (BN == 128)
      le = launch_conv2<128, 6, kF16>(npg, stream, tmA, tmS1, tmS2, tmB, tmO, a);
    else if (p->dtype == D3_DTYPE_BF16)
      le = launch_conv2<256, 5, kBF16>(npg, stream, tmA, tmS1, tmS2, tmB, tmO, a);
    else
      le = launch_conv2<256, 5, kF16>(npg, stream, tmA, tmS1, tmS2, tmB, tmO, a);
  } else if (use_t && p->dtype == D3_DTYPE_BF16)
    le = launch_conv_t<4, kBF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, tmR, a, use_strips, use_mc);
  else if (use_t)
    le = launch_conv_t<4, kF16>(grid, stream, tmA, tmS1, tmS2, tmB, tmO, tmR, a, use_strips, use_mc);
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
    if (p->dtype == D3_DTYPE_BF16) splitk_finalize_kernel<kBF16><<<fgrid, threads, 0, stream>>>(f);
    else splitk_finalize_kernel<kF16><<<fgrid, threads, 0, stream>>>(f);
  }
  D3_LAUNCH_CHECK();
  return D3_OK;
}
