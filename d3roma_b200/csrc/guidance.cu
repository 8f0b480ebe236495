// Stereo photometric guidance: analytic forward + gradient in ONE kernel.
//
// Replaces the autograd tape of FlowGuidance.grad's stereo branch (/root/reference/core/guidance.py:298-323:
// meshgrid/stack/grid_sample/5 avg-pools/reflection pads x (forward + backward), ~50 ATen launches) and SSIM.forward
// (:28-42).  Output = d(loss)/d(normalised disparity), loss = 0.95 * sum(SSIM_loss(left, warp)) + 0.05 * sum|left-warp|,
// warp = right image bilinearly sampled at x - disparity (border padding, align_corners=True),
// disparity = clip(min_d + (dn - shift) * (max_d - min_d), min_d, max_d)  (utils_d3roma/camera.py:599-602,778-780).
// The derivation is written out in SURVEY.md Appendix E; summary per channel with X = left, Y = warp, 3x3 windows q:
//   A1 = 2 mx my + C1, A2 = 2 sxy + C2, B1 = mx^2 + my^2 + C1, B2 = sx + sy + C2, n = A1 A2, d = B1 B2,
//   loss_q = clamp((1 - n/d)/2, 0, 1), and with f = n/d:
//     df/dmy = (2 mx (A2 - A1) d - n * 2 my (B2 - B1)) / d^2,  df/dEyy = -n B1 / d^2,  df/dExy = 2 A1 / d
//   dloss/dY(p) = (1/9) sum_{q in N(p)} w(q,p) * (-1/2) * mask_q * (df/dmy + 2 Y(p) df/dEyy + X(p) df/dExy)
//   where w(q,p) counts how often reflection padding maps a tap of window q onto p (2 at the image border rows/cols).
//   dY/ddn = -(R[x0+1] - R[x0]) * (max_d - min_d), zero where the disparity clip or the border clamp is active.
//
// Algorithmic bytes: (1 + 3 + 3 + 1) * B * H * W * 4.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

constexpr int TY = 16, TX = 32;           // output tile
constexpr int RY = TY + 4, RX = TX + 4;   // X / Y region (halo 2)
constexpr int GY = TY + 2, GX = TX + 2;   // window-coefficient region (halo 1)
constexpr int kGuidThreads = 256;

struct GuidArgs {
  const float* dn;     // [B,1,H,W]
  const float* left;   // [B,3,H,Wl]  (Wl = W + margin; the left image is cropped by `margin` columns)
  const float* right;  // [B,3,H,Wr]  (Wr = W + margin)
  float* grad;         // [B,1,H,W]
  int B, H, W, margin;
  float min_d, max_d, shift, w_ssim;
};

D3_DEVINL int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(kGuidThreads) guidance_grad_kernel(const GuidArgs a) {
  __shared__ float sX[RY][RX];
  __shared__ float sY[RY][RX];
  __shared__ float sDy[RY][RX];   // dY/ddn
  __shared__ float sG[3][GY][GX];  // per-window coefficients: d/dmy, d/dEyy, d/dExy (already * -1/2 * mask / 9)
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
  const int H = a.H, W = a.W, Wr = a.W + a.margin;
  const float C1 = 1e-4f, C2 = 9e-4f;
  const float scale = a.max_d - a.min_d;

  float acc[2] = {0.f, 0.f};  // each thread owns pixels tid and tid + 256 of the 16x32 tile
  for (int c = 0; c < 3; ++c) {
    const float* Lp = a.left + ((size_t)b * 3 + c) * H * Wr + a.margin;
    const float* Rp = a.right + ((size_t)b * 3 + c) * H * Wr;
    // ---- step 1: X, Y, dY/ddn on the halo-2 region (true image coordinates; outside the image -> 0)
    for (int i = threadIdx.x; i < RY * RX; i += kGuidThreads) {
      const int ry = i / RX, rx = i - ry * RX;
      const int y = y0 - 2 + ry, x = x0 - 2 + rx;
      float X = 0.f, Y = 0.f, dY = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float dnv = a.dn[((size_t)b * H + y) * W + x];
        const float raw = a.min_d + (dnv - a.shift) * scale;
        const float d = fminf(fmaxf(raw, a.min_d), a.max_d);
        const bool d_free = (raw >= a.min_d) && (raw <= a.max_d);
        // grid_sample(align_corners=True): the normalised coordinate round-trips to ix = x - d + margin
        float gx = ((float)x - d + (float)a.margin) / ((float)(Wr - 1) * 0.5f) - 1.0f;
        float ix = (gx + 1.0f) * 0.5f * (float)(Wr - 1);
        const bool x_free = (ix >= 0.0f) && (ix <= (float)(Wr - 1));
        ix = fminf(fmaxf(ix, 0.0f), (float)(Wr - 1));
        const int xl = (int)floorf(ix);
        const float fa = ix - (float)xl;
        const float r0 = Rp[(size_t)y * Wr + xl];
        const float r1 = (xl + 1 < Wr) ? Rp[(size_t)y * Wr + xl + 1] : 0.f;
        X = Lp[(size_t)y * Wr + x];
        Y = (1.0f - fa) * r0 + fa * r1;
        dY = (d_free && x_free) ? -(r1 - r0) * scale : 0.f;
      }
      sX[ry][rx] = X;
      sY[ry][rx] = Y;
      sDy[ry][rx] = dY;
    }
    __syncthreads();
    // ---- step 2: SSIM window coefficients on the halo-1 region
    for (int i = threadIdx.x; i < GY * GX; i += kGuidThreads) {
      const int gy = i / GX, gx_ = i - gy * GX;
      const int qy = y0 - 1 + gy, qx = x0 - 1 + gx_;
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int ry = reflect(qy + dy, H) - (y0 - 2);
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int rx = reflect(qx + dx, W) - (x0 - 2);
            const float X = sX[ry][rx], Y = sY[ry][rx];
            sx += X; sy += Y; sxx += X * X; syy += Y * Y; sxy += X * Y;
          }
        }
        const float inv9 = 1.0f / 9.0f;
        const float mx = sx * inv9, my = sy * inv9;
        const float vx = sxx * inv9 - mx * mx, vy = syy * inv9 - my * my, cxy = sxy * inv9 - mx * my;
        const float A1 = 2.f * mx * my + C1, A2 = 2.f * cxy + C2;
        const float B1 = mx * mx + my * my + C1, B2 = vx + vy + C2;
        const float n = A1 * A2, d = B1 * B2;
        const float s = (1.0f - n / d) * 0.5f;
        if (s >= 0.0f && s <= 1.0f) {
          const float inv_d = 1.0f / d;
          const float dfdmy = (2.f * mx * (A2 - A1) * d - n * 2.f * my * (B2 - B1)) * inv_d * inv_d;
          const float dfdeyy = -n * B1 * inv_d * inv_d;
          const float dfdexy = 2.f * A1 * inv_d;
          const float k = -0.5f * inv9;
          g0 = k * dfdmy; g1 = k * dfdeyy; g2 = k * dfdexy;
        }
      }
      sG[0][gy][gx_] = g0; sG[1][gy][gx_] = g1; sG[2][gy][gx_] = g2;
    }
    __syncthreads();
    // ---- step 3: gather the windows covering each pixel, chain through dY/ddn
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = threadIdx.x + u * kGuidThreads;
      const int ty = i / TX, tx = i - ty * TX;
      const int y = y0 + ty, x = x0 + tx;
      if (y < H && x < W) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int qy = y + dy;
          if (qy < 0 || qy >= H) continue;
          const float wy = ((qy == 0 && y == 1) || (qy == H - 1 && y == H - 2)) ? 2.f : 1.f;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int qx = x + dx;
            if (qx < 0 || qx >= W) continue;
            const float wx = ((qx == 0 && x == 1) || (qx == W - 1 && x == W - 2)) ? 2.f : 1.f;
            const float wgt = wy * wx;
            s0 += wgt * sG[0][ty + 1 + dy][tx + 1 + dx];
            s1 += wgt * sG[1][ty + 1 + dy][tx + 1 + dx];
            s2 += wgt * sG[2][ty + 1 + dy][tx + 1 + dx];
          }
        }
        const float X = sX[ty + 2][tx + 2], Y = sY[ty + 2][tx + 2];
        const float g_ssim = s0 + 2.f * Y * s1 + X * s2;
        const float diff = X - Y;
        const float g_l1 = (diff > 0.f) ? -1.f : ((diff < 0.f) ? 1.f : 0.f);
        acc[u] += (a.w_ssim * g_ssim + (1.0f - a.w_ssim) * g_l1) * sDy[ty + 2][tx + 2];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = threadIdx.x + u * kGuidThreads;
    const int ty = i / TX, tx = i - ty * TX;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) a.grad[((size_t)b * H + y) * W + x] = acc[u];
  }
}

}  // namespace
}  // namespace d3

extern "C" int d3_guidance_grad(const float* disp_norm, const float* left, const float* right, int B, int H, int W,
                                int margin, float min_disp, float max_disp, float shift, float w_ssim, float* grad,
                                void* stream) {
  using namespace d3;
  D3_REQUIRE(disp_norm && left && right && grad, "guidance: null pointer");
  D3_REQUIRE(H >= 2 && W >= 2 && margin >= 0, "guidance: bad shape");
  GuidArgs a{disp_norm, left, right, grad, B, H, W, margin, min_disp, max_disp, shift, w_ssim};
  dim3 grid(ceil_div(W, TX), ceil_div(H, TY), B);
  guidance_grad_kernel<<<grid, kGuidThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
