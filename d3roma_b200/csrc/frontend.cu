// Camera-frame front end of the robot / single-frame inference path (/root/reference/inference.py:99-213):
//   uint8 HWC frame (gray, RGB or RGBA)  -> cv2.resize(INTER_LINEAR) to the camera resolution -> HWC->CHW, .float()
//                                        -> (x / 255 - 0.5) * 2 -> replicate pad (InputPadder, utils.py:297-325)
//   fp32 depth map in metres             -> cv2.resize(INTER_NEAREST) -> disp = fxb / depth where depth > 0 else 0
//                                        -> Normalizer.normalize ("average" mode) -> replicate pad, + validity mask
// each as ONE kernel writing the padded fp32 NCHW tensors the pipeline consumes.  The bilinear resize reproduces
// OpenCV's 8-bit fixed-point arithmetic bit for bit (resize.cpp: 11-bit coefficients cvRound((1-f)*2048), cvRound(f*2048);
// horizontal pass in int; vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; column fractions are zeroed
// at the clamped borders, row fractions are not), pinned against cv2 itself in tests/golden/frontend.npz.
// Algorithmic bytes: Hs*Ws*C (uint8 in) + 3*Hp*Wp*4 (out) per frame; Hs*Ws*4 + (2*H*W + Hp*Wp)*4 per depth map.
#include "common.cuh"
#include "internal.h"

namespace d3 {
namespace {

struct LinCoef {
  int i0, i1;
  int a0, a1;
};

// OpenCV resize coefficient for destination index d: fx = (float)((d + 0.5) * scale - 0.5)  (scale in double)
D3_DEVINL LinCoef lin_coef(int d, double scale, int sn, bool zero_frac_at_border) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  LinCoef c;
  if (zero_frac_at_border) {   // x direction
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    c.i0 = s;
    c.i1 = min(s + 1, sn - 1);
  } else {                     // y direction: indices are clamped, the fraction is kept
    c.i0 = min(max(s, 0), sn - 1);
    c.i1 = min(max(s + 1, 0), sn - 1);
  }
  c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  c.a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  return c;
}

__global__ void __launch_bounds__(256) frame_to_input_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws,
                                                             int Cimg, int H, int W, int pt, int pl, int Hp, int Wp,
                                                             double sy, double sx, float* __restrict__ out) {
  const int64_t total = (int64_t)B * Hp * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xp = (int)(i % Wp);
    const int64_t r = i / Wp;
    const int yp = (int)(r % Hp);
    const int64_t b = r / Hp;
    const int y = min(max(yp - pt, 0), H - 1), x = min(max(xp - pl, 0), W - 1);   // replicate padding
    const uint8_t* img = src + (size_t)b * Hs * Ws * Cimg;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = (Cimg == 1) ? 0 : c;   // gray frames are tiled to 3 channels; RGBA drops alpha
      int v;
      if (Hs == H && Ws == W) {
        v = img[((size_t)y * Ws + x) * Cimg + cs];   // cv2.resize to the same size is the identity
      } else {
        const LinCoef cx = lin_coef(x, sx, Ws, true), cy = lin_coef(y, sy, Hs, false);
        const int s00 = img[((size_t)cy.i0 * Ws + cx.i0) * Cimg + cs], s01 = img[((size_t)cy.i0 * Ws + cx.i1) * Cimg + cs];
        const int s10 = img[((size_t)cy.i1 * Ws + cx.i0) * Cimg + cs], s11 = img[((size_t)cy.i1 * Ws + cx.i1) * Cimg + cs];
        const int S0 = s00 * cx.a0 + s01 * cx.a1, S1 = s10 * cx.a0 + s11 * cx.a1;
        v = (((cy.a0 * (S0 >> 4)) >> 16) + ((cy.a1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
      }
      // normalize_rgb_fn = (x / 255. - 0.5) * 2   (inference.py:193), fp32 operation order
      out[(((size_t)b * 3 + c) * Hp + yp) * Wp + xp] = __fmul_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.5f), 2.0f);
    }
  }
}

__global__ void __launch_bounds__(256) depth_to_input_kernel(const float* __restrict__ depth, int B, int Hs, int Ws,
                                                             int H, int W, int pt, int pl, int Hp, int Wp, double sy,
                                                             double sx, float fxb, float bound, float gamma, float t,
                                                             float s, float* __restrict__ raw_disp,
                                                             float* __restrict__ mask, float* __restrict__ norm_pad) {
  const int64_t total = (int64_t)B * Hp * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xp = (int)(i % Wp);
    const int64_t r = i / Wp;
    const int yp = (int)(r % Hp);
    const int64_t b = r / Hp;
    const int y = min(max(yp - pt, 0), H - 1), x = min(max(xp - pl, 0), W - 1);
    // cv2 INTER_NEAREST: source index = min(floor(dst * scale), size - 1)
    const int ys = min((int)floor((double)y * sy), Hs - 1), xs = min((int)floor((double)x * sx), Ws - 1);
    const float d = depth[((size_t)b * Hs + ys) * Ws + xs];
    const float disp = d > 0.f ? __fmul_rn(__frcp_rn(d), fxb) : 0.f;   // python_float / tensor = reciprocal * scalar
    float yv = __fdiv_rn(fminf(disp, bound), bound);
    if (gamma != 1.0f) yv = powf(yv, gamma);
    norm_pad[i] = __fmul_rn(__fsub_rn(yv, t), s);
    if (yp - pt == y && xp - pl == x) {   // interior pixel: also emit the unpadded maps
      const size_t o = ((size_t)b * H + y) * W + x;
      raw_disp[o] = disp;
      mask[o] = disp > 0.f ? 1.0f : 0.f;
    }
  }
}

inline int fe_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace
}  // namespace d3

extern "C" int d3_frame_to_input(const uint8_t* frames, int B, int Hs, int Ws, int Cimg, int H, int W, int pad_top,
                                 int pad_left, int Hp, int Wp, float* out, void* stream) {
  using namespace d3;
  D3_REQUIRE(frames && out, "frame_to_input: null pointer");
  D3_REQUIRE(Cimg == 1 || Cimg == 3 || Cimg == 4, "frame_to_input: frames must have 1, 3 or 4 channels (got %d)", Cimg);
  D3_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && pad_top >= 0 && pad_left >= 0 && Hp >= H + pad_top &&
                 Wp >= W + pad_left, "frame_to_input: bad geometry");
  frame_to_input_kernel<<<fe_blocks((int64_t)B * Hp * Wp), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, B, Hs, Ws, Cimg, H, W, pad_top, pad_left, Hp, Wp, 1.0 / ((double)H / Hs), 1.0 / ((double)W / Ws), out);  // cv::resize: scale = 1. / inv_scale
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_depth_to_input(const float* depth, int B, int Hs, int Ws, int H, int W, int pad_top, int pad_left,
                                 int Hp, int Wp, float fxb, float bound, float gamma, float t, float s,
                                 float* raw_disp, float* mask, float* norm_padded, void* stream) {
  using namespace d3;
  D3_REQUIRE(depth && raw_disp && mask && norm_padded, "depth_to_input: null pointer");
  D3_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && pad_top >= 0 && pad_left >= 0 && Hp >= H + pad_top &&
                 Wp >= W + pad_left && bound > 0.f, "depth_to_input: bad geometry");
  depth_to_input_kernel<<<fe_blocks((int64_t)B * Hp * Wp), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth, B, Hs, Ws, H, W, pad_top, pad_left, Hp, Wp, 1.0 / ((double)H / Hs), 1.0 / ((double)W / Ws), fxb, bound, gamma, t, s,
      raw_disp, mask, norm_padded);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
