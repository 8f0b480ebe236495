// Weight gradient of the 3x3 / 1x1 convolutions for sm_100a: a split-K tcgen05 GEMM whose reduction dimension is the
// PIXEL index, fed straight from the NHWC activations by TMA (no transposed copies).
//
//     dW[co][tap][ci] (+)= sum over (b, y, x) of  dy[b, y, x, co] * x[b, s*y + dh(tap), s*x + dw(tap), ci]
//
// GEMM view: M = 128 output channels, N = up to 256 (tap, ci) columns, K = pixels.  Both operands have the NON-reduced
// dimension contiguous in memory (channels), i.e. they are "MN-major" for the tensor core: one TMA box
// [64 channels, TW, TH, 1] (TW*TH = 64 pixels) lands as 64 rows of 128 B - exactly the SWIZZLE_128B MN-major canonical
// layout ((64 MN elements contiguous) x (8 K rows at 128 B), 8-row groups 1024 B apart); the instruction descriptor
// carries a_major = b_major = MN.  The box of the x operand is shifted by the filter tap, TMA zero-fills what falls
// outside the image (= the conv's zero padding), and a stride-2 conv reads x through the same parity-split 5-D view the
// forward kernel uses.  An N tile is four 64-channel boxes, each with its own (tap, channel block): layers with 128
// input channels put two taps side by side so that every MMA runs at N = 256 (an N = 128 tcgen05.mma is operand-port
// bound at half rate, see conv_igemm.cu).  Column u*64 + c of the accumulator IS column tap*Cin + cb*64 + c of the
// packed [Cout][taps*Cin] weight layout of the forward kernels, so the epilogue adds rows of the accumulator to rows of
// the fp32 gradient with vector reductions (red.global.add.v4.f32); the pixel range is split over the CTAs (split-K)
// so that a layer always fills the 148 SMs.
//
// 3x3 stride-1 layers run the HALO form (the kernel is L2-bandwidth bound: per 64-pixel K block the generic form pulls
// 48 KB through the L2 for 128 x 256 x 64 MACs, 14.7 GB for a 128 -> 128 layer at 16 x 384 x 640, which is what capped
// it at 620 TFLOP/s): the pixel tile is 8 x 8, so one tile row is exactly one 1024-byte swizzle atom, and the x operand
// of a (kw, channel block) pair is loaded ONCE as a 10-row halo block [64 ch, 8, 10] (10 KB); the three kh taps are the
// same block seen through descriptors whose 64-channel atoms start 0, 1 and 2 rows in (leading byte offset = 1024 B =
// one tile row: overlapping atoms of one N = 192 operand), i.e. ONE tcgen05.mma per K step covers three filter taps.
// A work item owns two such groups (384 fp32 accumulator columns): 36 KB of L2 traffic per K block for
// 128 x 384 x 64 MACs - half the bytes per MAC - and no half-empty N tile (128 -> 128: 3 items of 384 columns instead
// of 4 x 256 + 1 x 128).
//
// One persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue;
// generic form: 4-stage smem ring of {dy 16 KB, x 32 KB}, two 256-column fp32 accumulators in TMEM (the epilogue of one
// work item runs under the MMAs of the next); halo form: 6 stages of {dy 16 KB, 2 x 10 KB halo blocks}, one 384-column
// accumulator set.
//
// Replaces the weight half of aten::convolution_backward (cuDNN wgrad) in the training step of the reference
// (/root/reference/train.py:157-538, `accelerator.backward(loss)` at :455).
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "tensormap.h"

namespace d3 {
namespace {

constexpr int kWgThreads = 256;
constexpr int kWgStages = 4;
constexpr uint32_t kBoxBytes = 64 * 128;          // one [64 ch x 64 px] box
constexpr uint32_t kWgABytes = 2 * kBoxBytes;     // 128 output channels
constexpr uint32_t kWgBBytes = 4 * kBoxBytes;     // up to 256 (tap, ci) columns
constexpr uint32_t kWgStageBytes = kWgABytes + kWgBBytes;
// halo form: 8 x 8 pixel tiles, x as (8 + 2) rows of 8 pixels per (kw, channel block) group, two groups per stage
constexpr int kHaloStages = 6;
constexpr uint32_t kHaloBytes = 10 * 8 * 128;
constexpr uint32_t kHaloStageBytes = kWgABytes + 2 * kHaloBytes;

struct WgArgs {
  int B, Ho, Wo;
  int stride, pad;
  int cbPerTap;     // Cin / 64
  int U;            // ntaps * cbPerTap column boxes
  int nTilesN, nTilesM;
  int TH, TW, tilesH, tilesW, pixTiles;
  int S, tilesPer;  // split-K: S pixel ranges of tilesPer pixel tiles
  int numWork;
  signed char tap_dh[9], tap_dw[9];
  float* dw;
  long long dw_pitch;
  int sc_off2;      // stride 2: x pitch (element offset of the odd-column parity plane)
  int Cin;
  int G;            // halo form: 3 * cbPerTap (kw, channel block) groups, two per work item
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor: 64 MN elements (128 B) contiguous, 8 K rows 128 B apart,
// next 8-row K group `sbo` bytes on, next 64-element MN atom `lbo` bytes on (cute: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO))).
D3_DEVINL uint64_t make_sw128_mnmajor_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc_16bit_mn(uint32_t M, uint32_t N, bool bf16) {
  return make_idesc_16bit(M, N, bf16) | (1u << 15) | (1u << 16);   // A and B MN-major
}

D3_DEVINL void red_add_v4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)),
               "f"(__uint_as_float(b)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

struct WgWork {
  int m, n, t0, t1;
};
D3_DEVINL WgWork wg_decode(const WgArgs& a, int work) {
  WgWork w;
  w.m = work % a.nTilesM;
  int r = work / a.nTilesM;
  w.n = r % a.nTilesN;
  const int s = r / a.nTilesN;
  w.t0 = s * a.tilesPer;
  w.t1 = min(a.pixTiles, w.t0 + a.tilesPer);
  return w;
}

template <int DT, bool HALO>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgArgs a) {
  constexpr int kStages = HALO ? kHaloStages : kWgStages;
  constexpr uint32_t kStageBytes = HALO ? kHaloStageBytes : kWgStageBytes;
  constexpr int kAccs = HALO ? 1 : 2;   // halo form: one 384-column accumulator set; generic: two of 256 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
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
    tmem_alloc(tmem_ptr, 512);
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
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < a.numWork; work += gridDim.x) {
        const WgWork w = wg_decode(a, work);
        const int u0 = w.n * 4;
        const int cnt = HALO ? min(2, a.G - w.n * 2) : min(4, a.U - u0);
        for (int t = w.t0; t < w.t1; ++t) {
          int p = t;
          const int tw = p % a.tilesW;
          p /= a.tilesW;
          const int th = p % a.tilesH;
          const int b = p / a.tilesH;
          const int h0 = th * a.TH, w0 = tw * a.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + kWgABytes;
          mbar_arrive_expect_tx(&full_bar[stage], 2 * kBoxBytes + cnt * (HALO ? kHaloBytes : kBoxBytes));
          tma_load_4d(sA, &tmDY, &full_bar[stage], w.m * 128, w0, h0, b);
          tma_load_4d(sA + kBoxBytes, &tmDY, &full_bar[stage], w.m * 128 + 64, w0, h0, b);
          if constexpr (HALO) {
            for (int g = 0; g < cnt; ++g) {
              const int hg = w.n * 2 + g;              // group = (kw, channel block), kw-major
              const int kw = hg / a.cbPerTap;
              const int cb = hg - kw * a.cbPerTap;
              tma_load_4d(sB + g * kHaloBytes, &tmX, &full_bar[stage], cb * 64, w0 + kw - 1, h0 - 1, b);
            }
          } else
          for (int j = 0; j < cnt; ++j) {
            const int u = u0 + j;
            const int tap = u / a.cbPerTap;
            const int cb = u - tap * a.cbPerTap;
            if (a.stride == 1) {
              tma_load_4d(sB + j * kBoxBytes, &tmX, &full_bar[stage], cb * 64, w0 + a.tap_dw[tap], h0 + a.tap_dh[tap], b);
            } else {
              // input pixel (2*ho + dh, 2*wo + dw) in the parity-split 5-D view: row 2*hh + ph, hh = ho + floor(dh / 2)
              const int dh = a.tap_dh[tap], dw = a.tap_dw[tap];
              tma_load_5d(sB + j * kBoxBytes, &tmX, &full_bar[stage], (dw & 1) * a.sc_off2 + cb * 64, w0 + (dw >> 1),
                          dh & 1, h0 + (dh >> 1), b);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = blockIdx.x; work < a.numWork; work += gridDim.x) {
        const WgWork w = wg_decode(a, work);
        const int cnt = HALO ? min(2, a.G - w.n * 2) : min(4, a.U - w.n * 4);
        const uint32_t idesc = make_idesc_16bit_mn(128, HALO ? 192 : 64 * cnt, DT == kBF16);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int t = w.t0; t < w.t1; ++t) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint32_t sB = sA + kWgABytes;
          if constexpr (HALO) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // K step k = tile rows 2k, 2k + 1 (16 pixels); the three 64-channel atoms of the x operand (kh = 0, 1, 2)
              // start one tile row = one 1024-byte swizzle atom apart inside the halo block
              const uint64_t adesc = make_sw128_mnmajor_desc(sA + k * 2048, kBoxBytes, 1024);
              for (int g = 0; g < cnt; ++g) {
                const uint64_t bdesc = make_sw128_mnmajor_desc(sB + g * kHaloBytes + k * 2048, 1024, 1024);
                umma_bf16_ss(d_tmem + g * 192, adesc, bdesc, idesc, (t != w.t0 || k != 0) ? 1u : 0u);
              }
            }
          } else
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixels = two 8-row K groups = 2048 B further into every box
            // (validated on B200: leading byte offset = distance between the 64-channel boxes, stride byte offset =
            // distance between 8-pixel row groups; the swapped assignment gives wrong sums)
            const uint64_t adesc = make_sw128_mnmajor_desc(sA + k * 2048, kBoxBytes, 1024);
            const uint64_t bdesc = make_sw128_mnmajor_desc(sB + k * 2048, kBoxBytes, 1024);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (t != w.t0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == kAccs) acc = 0;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const uint32_t q = warp - 4;
    const uint32_t row = q * 32 + lane;   // accumulator lane == output channel inside the M tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < a.numWork; work += gridDim.x) {
      const WgWork w = wg_decode(a, work);
      const int cnt = HALO ? min(2, a.G - w.n * 2) : min(4, a.U - w.n * 4);
      float* drow = a.dw + (size_t)(w.m * 128 + row) * a.dw_pitch;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (HALO) {
#pragma unroll 1
        for (int c = 0; c < 6 * cnt; ++c) {   // 32-column chunks: group g = c / 6, kh = (c % 6) / 2, half = c & 1
          const int g = c / 6, kh = (c % 6) >> 1, half = c & 1;
          const int hg = w.n * 2 + g;
          const int kw = hg / a.cbPerTap;
          const int cb = hg - kw * a.cbPerTap;
          float* dst = drow + (size_t)(kh * 3 + kw) * a.Cin + cb * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + g * 192 + kh * 64 + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else {
        float* dst = drow + (size_t)w.n * 256;
#pragma unroll 1
        for (int c = 0; c < 2 * cnt; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * 256 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) red_add_v4(dst + c * 32 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == kAccs) acc = 0;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// pixel tile TH x TW = 64 that wastes the fewest rows
void choose_tile64(int Ho, int Wo, int* TH, int* TW) {
  double best = -1.0;
  for (int tw = 1; tw <= 64; tw *= 2) {
    const int th = 64 / tw;
    const double eff = double(Ho) * Wo / (double(ceil_div(Ho, th)) * th * ceil_div(Wo, tw) * tw);
    // prefer wide rows (longer contiguous runs per TMA row) on ties
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw <= 16)) {
      best = eff;
      *TH = th;
      *TW = tw;
    }
  }
}

}  // namespace
}  // namespace d3

extern "C" int d3_conv2d_wgrad(const d3_wgrad_params* p, void* stream_) {
  using namespace d3;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3_REQUIRE(p != nullptr && p->x && p->dy && p->dw, "wgrad: null pointer");
  D3_REQUIRE(p->dtype == D3_DTYPE_BF16 || p->dtype == D3_DTYPE_F16, "wgrad: 16-bit dtypes only (got %d)", p->dtype);
  D3_REQUIRE(p->ksize == 3 || p->ksize == 1, "wgrad: ksize must be 1 or 3 (got %d)", p->ksize);
  D3_REQUIRE(p->stride == 1 || (p->stride == 2 && p->ksize == 3), "wgrad: unsupported stride %d", p->stride);
  D3_REQUIRE(p->Cin > 0 && p->Cin % 64 == 0, "wgrad: Cin must be a multiple of 64 (got %d)", p->Cin);
  D3_REQUIRE(p->Cout > 0 && p->Cout % 128 == 0, "wgrad: Cout must be a multiple of 128 (got %d)", p->Cout);
  D3_REQUIRE(p->x_pitch % 8 == 0 && p->dy_pitch % 8 == 0 && p->dw_pitch % 4 == 0 && p->dw_col0 % 4 == 0,
             "wgrad: pitches must keep 16-byte alignment");
  D3_REQUIRE(p->stride == 1 || (p->H % 2 == 0 && p->W % 2 == 0), "wgrad: stride 2 needs even H, W");
  const int Ho = p->H / p->stride, Wo = p->W / p->stride;
  const int taps = p->ksize * p->ksize;
  D3_REQUIRE(p->dw_col0 + taps * p->Cin <= p->dw_pitch, "wgrad: column range exceeds the gradient row");

  WgArgs a{};
  a.B = p->B;
  a.Ho = Ho;
  a.Wo = Wo;
  a.stride = p->stride;
  a.pad = p->ksize / 2;
  a.cbPerTap = p->Cin / 64;
  a.U = taps * a.cbPerTap;
  a.Cin = p->Cin;
  a.G = 3 * a.cbPerTap;
  // halo form for every 3x3 stride-1 layer (D3_WGRAD_HALO=0 routes them to the generic per-tap form; both are
  // parity-tested)
  static int want_halo = -1;
  if (want_halo < 0) { const char* e = getenv("D3_WGRAD_HALO"); want_halo = e ? atoi(e) : 1; }
  const bool halo = want_halo && p->ksize == 3 && p->stride == 1;
  a.nTilesN = halo ? ceil_div(a.G, 2) : ceil_div(a.U, 4);
  a.nTilesM = p->Cout / 128;
  if (halo) a.TH = a.TW = 8;
  else choose_tile64(Ho, Wo, &a.TH, &a.TW);
  a.tilesH = ceil_div(Ho, a.TH);
  a.tilesW = ceil_div(Wo, a.TW);
  a.pixTiles = p->B * a.tilesH * a.tilesW;
  for (int t = 0; t < taps; ++t) {
    a.tap_dh[t] = (signed char)(t / p->ksize - a.pad);
    a.tap_dw[t] = (signed char)(t % p->ksize - a.pad);
  }
  // split-K: about 4 work items per SM, at least 16 pixel tiles (64 MMAs) per item
  const int mn = a.nTilesM * a.nTilesN;
  int S = ceil_div(num_sms() * 4, mn);
  if (S > a.pixTiles / 16) S = a.pixTiles / 16;
  if (S < 1) S = 1;
  a.tilesPer = ceil_div(a.pixTiles, S);
  a.S = ceil_div(a.pixTiles, a.tilesPer);
  a.numWork = mn * a.S;
  a.dw = p->dw + p->dw_col0;
  a.dw_pitch = p->dw_pitch;
  a.sc_off2 = p->x_pitch;

  CUtensorMap tmDY, tmX;
  int rc;
  const bool f16 = p->dtype == D3_DTYPE_F16;
  {
    uint64_t dims[4] = {(uint64_t)p->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->dy_pitch * 2, (uint64_t)Wo * p->dy_pitch * 2, (uint64_t)Ho * Wo * p->dy_pitch * 2};
    uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)a.TH, 1};
    if ((rc = make_tmap_16bit(f16, &tmDY, p->dy, 4, dims, str, box)))
      return set_error(D3_ERR_CUDA, "wgrad: tensor map dy encode failed (%d)", rc);
  }
  if (p->stride == 1) {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
    uint64_t str[3] = {(uint64_t)p->x_pitch * 2, (uint64_t)p->W * p->x_pitch * 2, (uint64_t)p->H * p->W * p->x_pitch * 2};
    uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)(halo ? a.TH + 2 : a.TH), 1};
    rc = make_tmap_16bit(f16, &tmX, p->x, 4, dims, str, box);
  } else {
    // parity-split view: (pw*pitch + c, w/2, h parity, h/2, b)
    uint64_t dims[5] = {(uint64_t)p->x_pitch + p->Cin, (uint64_t)p->W / 2, 2, (uint64_t)p->H / 2, (uint64_t)p->B};
    uint64_t str[4] = {(uint64_t)p->x_pitch * 4, (uint64_t)p->W * p->x_pitch * 2, (uint64_t)p->W * p->x_pitch * 4,
                       (uint64_t)p->H * p->W * p->x_pitch * 2};
    uint32_t box[5] = {64, (uint32_t)a.TW, 1, (uint32_t)a.TH, 1};
    rc = make_tmap_16bit(f16, &tmX, p->x, 5, dims, str, box);
  }
  if (rc) return set_error(D3_ERR_CUDA, "wgrad: tensor map x encode failed (%d)", rc);

  const size_t smem = (halo ? kHaloStages * kHaloStageBytes : kWgStages * kWgStageBytes) + 1024 + 256;
  const int grid = a.numWork < num_sms() ? a.numWork : num_sms();
  cudaError_t le;
#define D3_WG_LAUNCH(DT, H)                                                                                     \
  do {                                                                                                          \
    if ((le = ensure_dyn_smem(reinterpret_cast<const void*>(conv_wgrad_kernel<DT, H>), smem)) == cudaSuccess)   \
      le = launch_pdl(conv_wgrad_kernel<DT, H>, dim3(grid), dim3(kWgThreads), smem, stream, tmDY, tmX, a);      \
  } while (0)
  if (f16 && halo) D3_WG_LAUNCH(kF16, true);
  else if (f16) D3_WG_LAUNCH(kF16, false);
  else if (halo) D3_WG_LAUNCH(kBF16, true);
  else D3_WG_LAUNCH(kBF16, false);
#undef D3_WG_LAUNCH
  D3_CHECK_CUDA(le);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
