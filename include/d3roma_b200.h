/*
 * d3roma_b200 — C ABI of the B200-native disparity-denoising hot path.
 *
 * The reference (jie0530/d3roma) is pure Python: its "operator interface" for this path is
 *   - diffusers.UNet2DModel.forward              (constructed at train.py:788-810, called at
 *                                                 core/custom_pipelines.py:173)
 *   - MyDDIMScheduler.step / MyDDPMScheduler.step (core/scheduler_ddim.py:346-498,
 *                                                 core/scheduler_ddpm.py:403-526)
 *   - FlowGuidance.grad stereo branch            (core/guidance.py:293-331, SSIM :28-42)
 *   - Normalizer.normalize/denormalize, depth<->disparity
 *                                                (utils_d3roma/utils.py:171-273, inference.py:128-131,246-249)
 *   - InputPadder.pad/unpad                      (utils_d3roma/utils.py:297-325)
 * Each entry point below names the ATen op sequence / reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero code otherwise; d3_last_error() gives text.
 *   - all data pointers are DEVICE pointers unless the name ends in _host; `stream` is a cudaStream_t.
 *   - activations inside the UNet are bf16, NHWC, with an explicit pixel pitch (elements between
 *     consecutive pixels); pipeline-facing tensors (samples, images, disparities) are fp32 NCHW
 *     exactly as the reference passes them.
 *   - no C++ types, no torch types, no exceptions cross this boundary.
 */
#ifndef D3ROMA_B200_H
#define D3ROMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3_OK 0
#define D3_ERR_INVALID 1
#define D3_ERR_CUDA 2
#define D3_ERR_UNSUPPORTED 3

/* ---- library ------------------------------------------------------------------------------- */
const char* d3_last_error(void);
int d3_version(void);
/* Number of kernels launched by this library since the last d3_reset_launch_count() (host counter). */
int64_t d3_launch_count(void);
void d3_reset_launch_count(void);

/* ---- implicit-GEMM convolution (tcgen05 / TMEM / TMA) ---------------------------------------
 * Replaces aten::conv2d as invoked by ResnetBlock2D.conv1/conv2/conv_shortcut, Downsample2D.conv,
 * Upsample2D.conv and the attention Linear layers (as 1x1) — SURVEY.md §2.3 rows 1-3,6.
 * out[b,ho,wo,n] = bias[b?,n] + sum_{kh,kw,c} x[b, ho*s+kh-p, wo*s+kw-p, c] * w[n, (kh*k+kw)*Cin + c]
 *                + sum_c sc1[b,ho,wo,c] * w[n, k*k*Cin + c] + sum_c sc2[b,ho,wo,c] * w[n, k*k*Cin + C1 + c]
 *                + residual[b,ho,wo,n]
 * fp32 accumulation, bf16 output. Optionally accumulates per-(b, channel-quad) sum / sum-of-squares of
 * the ROUNDED outputs into `stats` (fp64) for the GroupNorm that follows. */
typedef struct d3_conv_params {
  const void* x;       /* bf16 NHWC [B,H,W,Cin], pixel pitch x_pitch                                */
  int32_t B, H, W;     /* INPUT spatial size                                                         */
  int32_t Cin;         /* multiple of 64                                                             */
  int32_t x_pitch;
  int32_t ksize;       /* 3 (padding 1) or 1 (padding 0)                                             */
  int32_t stride;      /* 1, or 2 (ksize 3 only; H and W even)                                       */
  const void* w;       /* bf16 [Cout][Ktot], Ktot = ksize*ksize*Cin + C1 + C2                        */
  int32_t Cout;        /* multiple of 128                                                            */
  const void* sc1;     /* optional bf16 NHWC [B,Ho,Wo,C1] folded 1x1 shortcut source (may be NULL)   */
  int32_t C1, sc1_pitch;
  const void* sc2;     /* optional second shortcut source (virtual concat)                           */
  int32_t C2, sc2_pitch;
  const float* bias;   /* fp32 [Cout] (bias_bstride 0) or [B][Cout] (bias_bstride = Cout)            */
  int32_t bias_bstride;
  const void* residual; /* optional bf16 NHWC [B,Ho,Wo,Cout] identity residual                       */
  int32_t res_pitch;
  void* out;           /* bf16 NHWC [B,Ho,Wo,Cout], pixel pitch out_pitch                            */
  int32_t out_pitch;
  double* stats;       /* optional fp64 [B][Cout/4][2], must be zero on entry; NULL to skip          */
} d3_conv_params;

int d3_conv2d_igemm(const d3_conv_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D3ROMA_B200_H */
