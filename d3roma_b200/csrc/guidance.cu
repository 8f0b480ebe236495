// Stereo photometric guidance: analytic forward + gradient (+ the guided x0 / x_{t-1} update) in ONE streaming kernel.
//
// Replaces the autograd tape of FlowGuidance.grad's stereo branch (/root/reference/core/guidance.py:298-323:
// meshgrid/stack/grid_sample/5 avg-pools/reflection pads x (forward + backward), ~50 ATen launches) and SSIM.forward
// (:28-42), and the recombination of the scheduler hook (core/scheduler_ddim.py:461-468, core/scheduler_ddpm.py:506-516).
//
//   disparity [px]  disp = clip(off + sc * sum_c x0_c, min_d, max_d)
//       reference grad() semantic (x0 = camera-normalised disparity, utils_d3roma/camera.py:599-602, 778-780):
//           sc = max_d - min_d, off = min_d - shift * sc, one channel;
//       guided sampling (x0 in Normalizer space, utils_d3roma/utils.py:260-273 "average", gamma = 1):
//           sc = bound / (s * C), off = t * bound  - the chain x0 -> pixels -> warp the reference's `denormer` hook stands for.
//   warp            Y = right image bilinearly sampled at x - disp (+ margin), border padding, align_corners=True
//   loss            w * sum(SSIM_loss(left, Y)) + (1 - w) * sum|left - Y|          (guidance.py:318-321)
//   output          g = d(loss)/d(x0_c) = sc * d(loss)/d(disp)  (identical for every channel c), and optionally
//                   x0' = x0 - step * g   and   out2 = ca * x0' + cb * aux1 (+ cc * aux2)   (the scheduler's recombination)
// The derivation is written out in SURVEY.md Appendix E; per channel with X = left, Y = warp, 3x3 windows q:
//   A1 = 2 mx my + C1, A2 = 2 sxy + C2, B1 = mx^2 + my^2 + C1, B2 = sx + sy + C2, n = A1 A2, d = B1 B2,
//   loss_q = clamp((1 - n/d)/2, 0, 1), and with f = n/d:
//     df/dmy = (2 mx (A2 - A1) d - n * 2 my (B2 - B1)) / d^2,  df/dEyy = -n B1 / d^2,  df/dExy = 2 A1 / d
//   dloss/dY(p) = (1/9) sum_{q in N(p)} w(q,p) * (-1/2) * mask_q * (df/dmy + 2 Y(p) df/dEyy + X(p) df/dExy)
//   where w(q,p) counts how often reflection padding maps a tap of window q onto p (2 at the image border rows/cols).
//   dY/ddisp = -(R[x0+1] - R[x0]), zero where the disparity clip or the border clamp is active.
//
// Execution: no shared memory, no block barriers.  One WARP owns a strip of 28 output columns (32 lanes = 28 + a halo of
// 2 on each side) and walks down kRows output rows; per input row every lane computes disparity / clip / floor ONCE and
// the three colour channels' X, Y, dY from it; the separable 3x3 sums travel between neighbouring columns through
// shuffles and between rows through a 3-deep register pipeline (row sums -> window coefficients -> gathered
// coefficients -> gradient, two rows behind the load).  Algorithmic bytes: (C + 3 + 3 + C) * B * H * W * 4.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int kStripW = 28;    // output columns per warp
constexpr int kRows = 32;      // output rows per warp
constexpr int kGuidWarps = 4;  // warps per block (independent)

struct GuidArgs {
  const float* x0;     // [B,C,H,W]
  const float* left;   // [B,3,H,Wr]  (Wr = W + margin; the left image is cropped by `margin` columns)
  const float* right;  // [B,3,H,Wr]
  float* grad;         // [B,1,H,W] or nullptr
  float* x0_out;       // [B,C,H,W] or nullptr
  float* out2;         // [B,C,H,W] or nullptr
  const float* aux1;   // [B,C,H,W] (required with out2)
  const float* aux2;   // [B,C,H,W] or nullptr
  int B, C, H, W, margin;
  float sc, off, min_d, max_d, w_ssim, step, ca, cb, cc;
  int stripsX, bands;
};

D3_DEVINL int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

struct Row5 {  // horizontal 3-sums of one image row for one channel
  float sx, sy, sxx, syy, sxy;
};

__global__ void __launch_bounds__(kGuidWarps * 32) guidance_stream_kernel(const GuidArgs a) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kGuidWarps + (threadIdx.x >> 5);
  const int total = a.B * a.bands * a.stripsX;
  if (wid >= total) return;
  const int strip = wid % a.stripsX;
  const int band = (wid / a.stripsX) % a.bands;
  const int b = wid / (a.stripsX * a.bands);
  const int H = a.H, W = a.W, Wr = a.W + a.margin;
  const int x = strip * kStripW - 2 + lane;          // image column of this lane (may be outside the image)
  const int xl_ = reflect_idx(x, W);                  // column whose data this lane carries (reflection padding)
  const bool x_in = (x >= 0) && (x < W);
  const bool x_ld = (xl_ >= 0) && (xl_ < W);          // false only for lanes far outside the image
  const bool x_out = x_in && lane >= 2 && lane < 2 + kStripW;
  // multiplicity of window column x+dx at pixel column x under reflection padding (0: window outside the image)
  const float wxm = (x - 1 < 0 || x - 1 >= W) ? 0.f : ((x - 1 == 0 && x == 1) ? 2.f : 1.f);
  const float wxp = (x + 1 < 0 || x + 1 >= W) ? 0.f : ((x + 1 == W - 1 && x == W - 2) ? 2.f : 1.f);
  const float C1 = 1e-4f, C2 = 9e-4f;
  const int y_first = band * kRows;
  const int y_end = min(H, y_first + kRows);
  const size_t plane = (size_t)H * W, planeR = (size_t)H * Wr;
  const float* x0b = a.x0 + (size_t)b * a.C * plane;
  const float* Lb = a.left + (size_t)b * 3 * planeR + a.margin;
  const float* Rb = a.right + (size_t)b * 3 * planeR;
  const float half_w = (float)(Wr - 1) * 0.5f;

  // register pipelines (index 0 = oldest row)
  Row5 rs[3][3];        // [channel][row slot]: horizontal sums of rows yr-2, yr-1, yr
  float gh[3][3][3];    // [channel][row slot][coef]: horizontally gathered window coefficients of rows qy-2, qy-1, qy
  float pX[3][3], pY[3][3], pD[3][3];  // [channel][row slot]: X, Y, dY/ddisp of rows yr-2, yr-1, yr
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      rs[c][s] = Row5{0.f, 0.f, 0.f, 0.f, 0.f};
      gh[c][s][0] = gh[c][s][1] = gh[c][s][2] = 0.f;
      pX[c][s] = pY[c][s] = pD[c][s] = 0.f;
    }

  // ---- software pipeline of the loads: the disparity of row yr+2 and the image gathers of row yr+1 (which depend on
  // the disparity of that row) are in flight while rows yr-2 .. yr are being processed - with 12 resident warps per SM
  // the two dependent load levels per row were fully exposed (ncu: issue 37 % busy, profiles/r02_ncu_full_summary.csv)
  struct RowLoads {   // what stage A needs of one input row
    float r0[3], r1[3], X[3], fa, dsel;
  };
  auto load_xsum = [&](int yrow) -> float {   // sum_c x0 of image row reflect(yrow) at this lane's column (0 if none)
    const int yl = reflect_idx(yrow, H);
    float ssum = 0.f;
    if (yl >= 0 && yl < H && x_ld)
      for (int c = 0; c < a.C; ++c) ssum += __ldg(x0b + (size_t)c * plane + (size_t)yl * W + xl_);
    return ssum;
  };
  auto issue_row = [&](int yrow, float ssum, RowLoads& L) {
    const int yl = reflect_idx(yrow, H);
    L.fa = 0.f;
    L.dsel = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) L.r0[c] = L.r1[c] = L.X[c] = 0.f;
    if (yl >= 0 && yl < H && x_ld) {
      const float raw = fmaf(a.sc, ssum, a.off);
      const float d = fminf(fmaxf(raw, a.min_d), a.max_d);
      const bool d_free = (raw >= a.min_d) && (raw <= a.max_d);
      // grid_sample(align_corners=True): the normalised coordinate round-trips to ix = x - d + margin
      const float gx = ((float)xl_ - d + (float)a.margin) / half_w - 1.0f;
      float ix = (gx + 1.0f) * 0.5f * (float)(Wr - 1);
      const bool x_free = (ix >= 0.0f) && (ix <= (float)(Wr - 1));
      ix = fminf(fmaxf(ix, 0.0f), (float)(Wr - 1));
      const int xi = (int)floorf(ix);
      L.fa = ix - (float)xi;
      const bool has1 = xi + 1 < Wr;
      L.dsel = (d_free && x_free) ? -1.f : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* Rp = Rb + (size_t)c * planeR + (size_t)yl * Wr;
        L.r0[c] = __ldg(Rp + xi);
        L.r1[c] = has1 ? __ldg(Rp + xi + 1) : 0.f;
        L.X[c] = __ldg(Lb + (size_t)c * planeR + (size_t)yl * Wr + xl_);
      }
    }
  };
  RowLoads cur, nxt;
  issue_row(y_first - 2, load_xsum(y_first - 2), cur);
  float xsum_next = load_xsum(y_first - 1);

  for (int yr = y_first - 2; yr < y_end + 2; ++yr) {
    // ---------------- stage A: X, Y, dY of input row yr (reflected at the image border) from the loads issued one
    // iteration ago; then put the next row's gathers and the disparity of the row after it in flight
    float Xc[3], Yc[3], Dc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Xc[c] = cur.X[c];
      Yc[c] = (1.0f - cur.fa) * cur.r0[c] + cur.fa * cur.r1[c];
      Dc[c] = cur.dsel * (cur.r1[c] - cur.r0[c]);
    }
    if (yr + 1 < y_end + 2) issue_row(yr + 1, xsum_next, nxt);
    if (yr + 2 < y_end + 2) xsum_next = load_xsum(yr + 2);
    // ---------------- stage B: horizontal sums of row yr, window coefficients of row qy = yr - 1
    const int qy = yr - 1;
    const bool q_in = (qy >= 0) && (qy < H) && x_in;
    float gnew[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float Xm = __shfl_up_sync(0xffffffffu, Xc[c], 1), Xp = __shfl_down_sync(0xffffffffu, Xc[c], 1);
      const float Ym = __shfl_up_sync(0xffffffffu, Yc[c], 1), Yp = __shfl_down_sync(0xffffffffu, Yc[c], 1);
      rs[c][0] = rs[c][1];
      rs[c][1] = rs[c][2];
      Row5 r;
      r.sx = Xm + Xc[c] + Xp;
      r.sy = Ym + Yc[c] + Yp;
      r.sxx = fmaf(Xm, Xm, fmaf(Xc[c], Xc[c], Xp * Xp));
      r.syy = fmaf(Ym, Ym, fmaf(Yc[c], Yc[c], Yp * Yp));
      r.sxy = fmaf(Xm, Ym, fmaf(Xc[c], Yc[c], Xp * Yp));
      rs[c][2] = r;
      pX[c][0] = pX[c][1]; pX[c][1] = pX[c][2]; pX[c][2] = Xc[c];
      pY[c][0] = pY[c][1]; pY[c][1] = pY[c][2]; pY[c][2] = Yc[c];
      pD[c][0] = pD[c][1]; pD[c][1] = pD[c][2]; pD[c][2] = Dc[c];
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (q_in) {
        const float inv9 = 1.0f / 9.0f;
        const float mx = (rs[c][0].sx + rs[c][1].sx + rs[c][2].sx) * inv9;
        const float my = (rs[c][0].sy + rs[c][1].sy + rs[c][2].sy) * inv9;
        const float vx = (rs[c][0].sxx + rs[c][1].sxx + rs[c][2].sxx) * inv9 - mx * mx;
        const float vy = (rs[c][0].syy + rs[c][1].syy + rs[c][2].syy) * inv9 - my * my;
        const float cxy = (rs[c][0].sxy + rs[c][1].sxy + rs[c][2].sxy) * inv9 - mx * my;
        const float A1 = 2.f * mx * my + C1, A2 = 2.f * cxy + C2;
        const float B1 = mx * mx + my * my + C1, B2 = vx + vy + C2;
        const float n = A1 * A2, dd = B1 * B2;
        const float inv_d = 1.0f / dd;
        const float s = (1.0f - n * inv_d) * 0.5f;
        if (s >= 0.0f && s <= 1.0f) {
          const float k = -0.5f * inv9;
          g0 = k * (2.f * mx * (A2 - A1) * dd - n * 2.f * my * (B2 - B1)) * inv_d * inv_d;
          g1 = k * (-n * B1 * inv_d * inv_d);
          g2 = k * (2.f * A1 * inv_d);
        }
      }
      gnew[c][0] = g0; gnew[c][1] = g1; gnew[c][2] = g2;
    }
    // ---------------- stage C: horizontal gather of the coefficients of row qy, output row y = yr - 2
    const int y = yr - 2;
    const bool y_out = (y >= y_first) && (y < y_end);
    // multiplicity of window row y+dy at pixel row y (0: outside)
    const float wym = (y - 1 < 0 || y - 1 >= H) ? 0.f : ((y - 1 == 0 && y == 1) ? 2.f : 1.f);
    const float wyp = (y + 1 < 0 || y + 1 >= H) ? 0.f : ((y + 1 == H - 1 && y == H - 2) ? 2.f : 1.f);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float gm = __shfl_up_sync(0xffffffffu, gnew[c][k], 1), gp = __shfl_down_sync(0xffffffffu, gnew[c][k], 1);
        gh[c][0][k] = gh[c][1][k];
        gh[c][1][k] = gh[c][2][k];
        gh[c][2][k] = fmaf(wxm, gm, fmaf(wxp, gp, gnew[c][k]));
      }
      // rows: gh[c][0] = window row y-1, gh[c][1] = y, gh[c][2] = y+1; pX/pY/pD[c][0] = pixel row y
      const float s0 = fmaf(wym, gh[c][0][0], fmaf(wyp, gh[c][2][0], gh[c][1][0]));
      const float s1 = fmaf(wym, gh[c][0][1], fmaf(wyp, gh[c][2][1], gh[c][1][1]));
      const float s2 = fmaf(wym, gh[c][0][2], fmaf(wyp, gh[c][2][2], gh[c][1][2]));
      const float X = pX[c][0], Y = pY[c][0];
      const float g_ssim = s0 + 2.f * Y * s1 + X * s2;
      const float diff = X - Y;
      const float g_l1 = (diff > 0.f) ? -1.f : ((diff < 0.f) ? 1.f : 0.f);
      acc = fmaf(a.w_ssim * g_ssim + (1.0f - a.w_ssim) * g_l1, pD[c][0], acc);
    }
    if (y_out && x_out) {
      const float g = a.sc * acc;
      const size_t pix = (size_t)y * W + x;
      if (a.grad != nullptr) a.grad[(size_t)b * plane + pix] = g;
      if (a.x0_out != nullptr || a.out2 != nullptr) {
        const float wg = __fmul_rn(a.step, g);
        for (int c = 0; c < a.C; ++c) {
          const size_t o = ((size_t)b * a.C + c) * plane + pix;
          const float xp = __fsub_rn(a.x0[o], wg);        // x0' = x0 - step * g      (d3_lincomb rounding order)
          if (a.x0_out != nullptr) a.x0_out[o] = xp;
          if (a.out2 != nullptr) {
            float v = __fadd_rn(__fmul_rn(a.ca, xp), __fmul_rn(a.cb, a.aux1[o]));
            if (a.aux2 != nullptr) v = __fadd_rn(v, __fmul_rn(a.cc, a.aux2[o]));
            a.out2[o] = v;
          }
        }
      }
    }
    cur = nxt;
  }
}

}  // namespace
}  // namespace d3

extern "C" int d3_guidance_step(const d3_guidance_params* p, void* stream) {
  using namespace d3;
  D3_REQUIRE(p != nullptr, "guidance: null params");
  D3_REQUIRE(p->x0 && p->left && p->right, "guidance: null pointer");
  D3_REQUIRE(p->grad || p->x0_out || p->out2, "guidance: no output requested");
  D3_REQUIRE(p->H >= 2 && p->W >= 2 && p->margin >= 0 && p->B >= 1, "guidance: bad shape");
  D3_REQUIRE(p->C == 1 || p->C == 3, "guidance: C must be 1 or 3 (got %d)", p->C);
  D3_REQUIRE(!p->out2 || p->aux1, "guidance: out2 needs aux1");
  GuidArgs a{};
  a.x0 = p->x0; a.left = p->left; a.right = p->right;
  a.grad = p->grad; a.x0_out = p->x0_out; a.out2 = p->out2; a.aux1 = p->aux1; a.aux2 = p->aux2;
  a.B = p->B; a.C = p->C; a.H = p->H; a.W = p->W; a.margin = p->margin;
  a.sc = p->disp_scale; a.off = p->disp_offset; a.min_d = p->min_disp; a.max_d = p->max_disp;
  a.w_ssim = p->w_ssim; a.step = p->step; a.ca = p->ca; a.cb = p->cb; a.cc = p->cc;
  a.stripsX = ceil_div(p->W, kStripW);
  a.bands = ceil_div(p->H, kRows);
  const long long warps = (long long)p->B * a.bands * a.stripsX;
  D3_REQUIRE(warps < (1ll << 30), "guidance: problem too large");
  D3_CHECK_CUDA(launch_pdl(guidance_stream_kernel, dim3((unsigned)ceil_div((int)warps, kGuidWarps)), dim3(kGuidWarps * 32), 0, static_cast<cudaStream_t>(stream), a));
  D3_LAUNCH_CHECK();
  return D3_OK;
}

// reference grad() semantic: x0 is the camera-normalised disparity (one channel), output d(loss)/d(x0)
extern "C" int d3_guidance_grad(const float* disp_norm, const float* left, const float* right, int B, int H, int W,
                                int margin, float min_disp, float max_disp, float shift, float w_ssim, float* grad,
                                void* stream) {
  d3_guidance_params p{};
  p.x0 = disp_norm; p.left = left; p.right = right; p.grad = grad;
  p.B = B; p.C = 1; p.H = H; p.W = W; p.margin = margin;
  p.disp_scale = max_disp - min_disp;
  p.disp_offset = min_disp - shift * (max_disp - min_disp);
  p.min_disp = min_disp; p.max_disp = max_disp; p.w_ssim = w_ssim;
  return d3_guidance_step(&p, stream);
}
