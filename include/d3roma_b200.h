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

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3_OK 0
#define D3_ERR_INVALID 1
#define D3_ERR_CUDA 2
#define D3_ERR_UNSUPPORTED 3

/* 16-bit storage / tensor-core operand type of the UNet activations and packed weights. Both run tcgen05
 * kind::f16 at the same rate with fp32 accumulation; F16 carries 3 more mantissa bits (parity). */
#define D3_DTYPE_BF16 0
#define D3_DTYPE_F16 1
/* fp32 storage + CUDA-core FFMA arithmetic: the reference-precision path for the <= 1e-4 tolerance configuration
 * (BASELINE configs[3]); accepted by every entry point that takes a dtype, ~40x slower than the tensor-core path. */
#define D3_DTYPE_F32 2

/* ---- library ------------------------------------------------------------------------------- */
const char* d3_last_error(void);
int d3_version(void);
/* Number of kernels launched by this library since the last d3_reset_launch_count() (host counter). */
int64_t d3_launch_count(void);
void d3_reset_launch_count(void);
/* Programmatic dependent launch for the kernels of the denoising loop (the next kernel's blocks start their prologue
 * while the previous kernel drains): measured +2 % at batch 1 (latency case), -2 % at batch 16 on B200, so the UNet turns
 * it on for small batches only.  Returns the previous setting; D3_PDL=1 in the environment sets the initial value. */
int d3_set_pdl(int on);

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
  const float* bias;   /* fp32 [Cout] (bias_bstride 0) or [B][Cout] (bias_bstride = row stride in elements)            */
  int32_t bias_bstride;
  const void* residual; /* optional bf16 NHWC [B,Ho,Wo,Cout] identity residual                       */
  int32_t res_pitch;
  void* out;           /* bf16 NHWC [B,Ho,Wo,Cout], pixel pitch out_pitch                            */
  int32_t out_pitch;
  double* stats;       /* optional fp64 [B][Cout/4][2], must be zero on entry; NULL to skip          */
  int32_t dtype;       /* D3_DTYPE_BF16 / D3_DTYPE_F16: element type of x, w, sc1, sc2, residual, out        */
  int32_t up2x_phase;  /* 0 = plain conv. 1..4 = sub-pixel phase of "nearest-2x upsample + 3x3 conv" (Upsample2D):
                          x is the LOW-res tensor [B,H,W,Cin], out the HIGH-res tensor [B,2H,2W,Cout]; this launch
                          writes output pixels (2i+py, 2j+px), (py,px) = ((phase-1)/2, (phase-1)%2), from the 2x2
                          input neighbourhood with pre-summed weights w[Cout][4*Cin]. 4 launches = the whole layer
                          at 4/9 of the FLOPs and without materialising the upsampled tensor.                   */
  void* workspace;     /* optional fp32 scratch for split-K (small layers / small batches: too few tiles to fill the
                          SMs, so the K loop is split across CTAs and a second kernel sums the partials); size from
                          d3_conv2d_workspace_bytes(); NULL = never split                                           */
  int64_t workspace_bytes;
  int32_t Cout_valid;  /* 0 = Cout.  Otherwise the number of channels that exist in `out` / `residual` (a multiple of 64): the weight / bias rows [Cout_valid, Cout) are zero padding up to the
                          128-channel MMA tile (the 320-channel layers of the latent UNet run with Cout = 384) and are
                          clipped by the TMA store.  No statistics, no split-K in this mode.                          */
  int32_t asym_pad;    /* stride-2 3x3 only: 1 = input padded by (0,1,0,1) instead of 1 on every side - Downsample2D
                          with padding=0 of the VAE encoder (F.pad(x, (0,1,0,1)) + conv stride 2)                     */
} d3_conv_params;

/* bytes of split-K scratch this call would like (0 when it runs unsplit) */
size_t d3_conv2d_workspace_bytes(const d3_conv_params* p);
int d3_conv2d_igemm(const d3_conv_params* p, void* stream);

/* ---- GroupNorm(32 groups) + optional SiLU over up to two concatenated sources ------------------
 * Replaces aten::native_group_norm + aten::silu (+ the torch.cat of the up blocks) - SURVEY.md §2.3 rows 4,5,10.
 * out[b,p,c] = act(gamma[c] * (x[b,p,c] - mean[b,g]) * rstd[b,g] + beta[c]),  x = cat(x1, x2) along channels.
 * mean/rstd come from the producers' quad statistics (stats1 [B][C1/4][2], stats2 [B][C2/4][2], fp64). */
int d3_groupnorm_silu(const void* x1, int C1, int pitch1, const double* stats1, const void* x2, int C2, int pitch2,
                      const double* stats2, const float* gamma, const float* beta, int B, int HW, int groups,
                      float eps, int apply_silu, void* out, int out_pitch, int dtype, void* stream);

/* ---- the same pass with TWO outputs from one read of a single-source tensor: out_a = act_a(GN_a(x)) and
 * out_b = act_b(GN_b(x)) with separate affine parameters and groupings (groups_* groups over the C channels of x).
 * Used for skip tensors: the down path normalises x for its next conv (a) and, in the same pass, writes the up path's
 * normalised copy (b) straight into the channel range of the up-block's concatenated input (out_b pre-offset,
 * pitch_b = C_x + C_skip), so the up-block GroupNorm no longer re-reads the skip (ResnetBlock2D.norm1 over
 * torch.cat([hidden, skip]) in diffusers' UpBlock2D).  16-bit dtypes only. */
int d3_groupnorm_silu_dual(const void* x, int C, int pitch, const double* stats, int B, int HW, float eps,
                           const float* gamma_a, const float* beta_a, int groups_a, int silu_a, void* out_a,
                           int pitch_a, const float* gamma_b, const float* beta_b, int groups_b, int silu_b,
                           void* out_b, int pitch_b, int dtype, void* stream);

/* ---- nearest-neighbour 2x upsample (F.interpolate(scale_factor=2, mode="nearest") of Upsample2D) */
int d3_upsample_nearest2x(const void* in, int B, int H, int W, int C, int in_pitch, void* out, int out_pitch,
                          void* stream);

/* ---- conv_in front end: torch.cat([sample, cond], 1) + cast + NCHW->NHWC + 3x3 im2col with zero padding.
 * sample fp32 NCHW [B,Cs,H,W], cond fp32 NCHW [B,Cc,H,W] -> bf16 [B,H,W,128], k = tap*(Cs+Cc) + c, zero-padded.
 * Replaces custom_pipelines.py:158-171 (the per-step torch.cat) and the im2col of conv_in. */
int d3_im2col_input(const float* sample, int Cs, const float* cond, int Cc, int B, int H, int W, void* out,
                    int dtype, void* stream);

/* ---- conv_in in split form: conv_in(cat[x, cond]) = conv3x3(x; W[:, :Cs]) + {conv3x3(cond; W[:, Cs:]) + bias}.
 * The braces are constant over one sampling loop (custom_pipelines.py:158-171 re-concatenates the same conditioning
 * every step), so the caller computes them once (out32 mode: planes = cond, base = NULL, bias given -> fp32 NHWC
 * [B,H,W,128]) and every step adds the taps of the noisy sample (out16 mode: planes = sample, base = that tensor ->
 * 16-bit NHWC [B,H,W,128] + GroupNorm quad statistics [B][32][2]).  planes fp32 NCHW [B,C,H,W]; wT fp32
 * [C*9][128] with k = c*9 + (kh*3+kw).  Exactly one of out32 / out16 is non-NULL.  fp32 arithmetic throughout.
 * The per-loop term may also be kept in fp16 (computed with out16 + dtype = D3_DTYPE_F16, no stats; passed back as
 * base with base_dtype = D3_DTYPE_F16): per-step traffic (2 + 2) * 128 * P bytes instead of (4 + 2) * 128 * P. */
int d3_conv_in_direct(const float* planes, int C, const float* wT, const void* base, int base_dtype,
                      const float* bias, int B, int H, int W, float* out32, void* out16, int dtype, double* stats,
                      void* stream);

/* ---- conv_norm_out + SiLU + conv_out fused: bf16 NHWC [B,H,W,128] -> fp32 NCHW [B,dc,H,W] (dc = 1 or 3).
 * w fp32 [dc][9][128] (tap-major), stats = quad statistics of x ([B][32][2]).  16-bit dtypes run as a streaming
 * per-pixel GEMM into `workspace` (9*dc fp32 tap maps, d3_conv_out_workspace_bytes(B,H,W,dc) bytes, caller-owned
 * scratch) followed by a 9-neighbour gather; D3_DTYPE_F32 ignores the workspace. */
size_t d3_conv_out_workspace_bytes(int B, int H, int W, int dc);
int d3_conv_out_fused(const void* x, int pitch, const double* stats, const float* gamma, const float* beta,
                      const float* w, const float* bias, int B, int H, int W, int dc, float eps, float* out,
                      void* workspace, int dtype, void* stream);

/* ---- timestep embedding (Timesteps(128) -> Linear -> SiLU -> Linear -> SiLU) and all ResNet projections.
 * t fp32 [nt] (nt = 1 for the sampling loop, = B for per-sample timesteps). Weights are passed transposed:
 * w1t [128][512], w2t [512][512]. Output silu_emb [nt][512]. */
int d3_time_embed(const float* t, int nt, const float* w1t, const float* b1, const float* w2t, const float* b2,
                  float* silu_emb_out, void* stream);
/* out[n][r] = bias_all[r] + w_all[r,:] . silu_emb[n,:]; w_all fp32 [rows][512] = all time_emb_proj weights stacked,
 * bias_all = time_emb_proj.bias + conv1.bias stacked. */
int d3_temb_project_all(const float* silu_emb, int nt, const float* w_all, const float* bias_all, int rows,
                        float* out, void* stream);

/* ---- self-attention core, head_dim 8: qkv bf16 [B,N,>=3C] (q | k | v at channel offsets 0, C, 2C) -> [B,N,C] */
int d3_attention_hd8(const void* qkv, int pitch, int B, int N, int C, void* out, int out_pitch, int dtype,
                     void* stream);

/* ---- fused scheduler updates -----------------------------------------------------------------
 * prediction_type: 0 epsilon, 1 sample, 2 v_prediction.  clip_mode: 0 none, 1 clamp(+-clip_range),
 * 2 "thresholding" (== clamp(+-1)/1 because sample_max_value defaults to 1.0). */
typedef struct d3_ddim_coeffs {
  int32_t prediction_type;
  float sqrt_alpha_t;    /* alpha_prod_t ** 0.5                      (scheduler_ddim.py:416,424) */
  float sqrt_beta_t;     /* (1 - alpha_prod_t) ** 0.5                                          */
  float sqrt_alpha_prev; /* alpha_prod_t_prev ** 0.5                 (:459)                    */
  float dir_coeff;       /* (1 - alpha_prod_t_prev - std_dev_t**2) ** 0.5 (:456)               */
  float std_dev;         /* eta * variance ** 0.5                    (:448-449)                */
  int32_t clip_mode;
  float clip_range;
  int32_t use_clipped_model_output;
} d3_ddim_coeffs;

typedef struct d3_ddpm_coeffs {
  int32_t prediction_type;
  float sqrt_alpha_t;
  float sqrt_beta_t;
  float x0_coeff;        /* pred_original_sample_coeff               (scheduler_ddpm.py:482)   */
  float xt_coeff;        /* current_sample_coeff                     (:483)                    */
  float sigma;           /* multiplier of the N(0,1) draw for t > 0  (:496-504)                */
  int32_t clip_mode;
  float clip_range;
} d3_ddpm_coeffs;

/* fp32, n elements. variance_noise may be NULL (eta == 0 / t == 0). prev_mean (x_{t-1} before noise) and dir_out
 * (the "direction pointing to x_t" term) are optional outputs. Replaces scheduler_ddim.py:423-489. */
int d3_sched_ddim_step(const float* model_output, const float* sample, const float* variance_noise,
                       const d3_ddim_coeffs* c, int64_t n, float* prev_sample, float* pred_original,
                       float* prev_mean, float* dir_out, void* stream);
/* Replaces scheduler_ddpm.py:460-518. */
int d3_sched_ddpm_step(const float* model_output, const float* sample, const float* variance_noise,
                       const d3_ddpm_coeffs* c, int64_t n, float* prev_sample, float* pred_original,
                       float* prev_mean, void* stream);
/* conv_norm_out + SiLU + conv_out (d3_conv_out_fused) with the DDIM update (eta = 0) applied to the model output in
 * the same gather pass: sample = x_t fp32 [B,dc,H,W]; writes prev_sample = x_{t-1} and pred_original = x0 exactly as
 * d3_sched_ddim_step would (bit-identical), model_out optional (may be NULL).  16-bit dtypes only.  Replaces
 * custom_pipelines.py:173-175 (unet(...).sample followed by scheduler.step) for the unguided loop. */
int d3_conv_out_fused_ddim(const void* x, int pitch, const double* stats, const float* gamma, const float* beta,
                           const float* w, const float* bias, int B, int H, int W, int dc, float eps, float* model_out,
                           void* workspace, int dtype, const float* sample, const d3_ddim_coeffs* c, float* prev_sample,
                           float* pred_original, void* stream);
/* out = a*x (+ b*y) (+ c*z), each product and sum rounded separately (recombination after a guidance
 * perturbation of x0: scheduler_ddim.py:468, scheduler_ddpm.py:512). y, z may be NULL. */
int d3_lincomb(float a, const float* x, float b, const float* y, float c, const float* z, int64_t n, float* out,
               void* stream);

/* ---- depth map -> colour-camera frame (DepthCamera.transform_depth_to_rgb_frame and
 * transform_cropped_depth_to_rgb_frame, utils_d3roma/camera.py:610-701): back-project with the inverse depth intrinsics
 * (row-major 3x3, float64), apply T_cl (row-major 3x4), project with the colour intrinsics, write the source depth at the
 * rounded target pixel of a canvas_h x canvas_w map and return the window [crop_y, crop_y+out_h) x [crop_x, crop_x+out_w).
 * numpy's collision order is reproduced exactly: within one fill the last source pixel in row-major order wins;
 * aggressive_fill runs the four (floor|ceil) corner fills in the reference's order, later fills overwriting earlier
 * ones.  depth fp32 [H,W] (device); the three matrices are HOST pointers; workspace = device scratch of
 * d3_depth_to_rgb_frame_workspace_bytes(canvas_h, canvas_w) bytes. */
size_t d3_depth_to_rgb_frame_workspace_bytes(int canvas_h, int canvas_w);
int d3_depth_to_rgb_frame(const float* depth, int H, int W, const double* Kinv_depth, const double* T_cl,
                          const double* K_color, int canvas_h, int canvas_w, int aggressive_fill, int crop_y,
                          int crop_x, int out_h, int out_w, void* workspace, float* out, void* stream);

/* ---- stereo photometric guidance (FlowGuidance.grad stereo branch, core/guidance.py:298-323 + SSIM :28-42) and the
 * guided update of the scheduler hook (core/scheduler_ddim.py:461-468, core/scheduler_ddpm.py:506-516), one kernel.
 *   disparity [px] = clip(disp_offset + disp_scale * sum_c x0[b,c,y,x], min_disp, max_disp)
 *   warp           = right sampled at x - disparity (+ margin), bilinear, border padding, align_corners
 *   loss           = w_ssim * sum SSIM_loss(left, warp) + (1 - w_ssim) * sum |left - warp|
 *   g[b,0,y,x]     = d(loss)/d(x0[b,c,y,x]) = disp_scale * d(loss)/d(disparity)        (the same for every c)
 *   x0_out         = x0 - step * g                                   (FlowGuidance.perturb, core/guidance.py:214)
 *   out2           = ca * x0_out + cb * aux1 (+ cc * aux2)           (each product / sum rounded separately)
 * The affine map x0 -> pixels is what the reference's `denormer` hook argument stands for (Normalizer.denormalize,
 * utils_d3roma/utils.py:260-273, "average" mode, gamma 1: disp_scale = bound / (s * C), disp_offset = t * bound); the
 * reference's own grad() takes camera-normalised disparity (utils_d3roma/camera.py:599-602, 771-780:
 * disp_scale = max - min, disp_offset = min - shift * (max - min)) - d3_guidance_grad below.
 * x0, x0_out, out2, aux1, aux2 fp32 [B,C,H,W] (C = 1 or 3); grad fp32 [B,1,H,W]; left, right fp32 [B,3,H,W+margin].
 * Any of grad / x0_out / out2 may be NULL (at least one is required). */
typedef struct d3_guidance_params {
  const float* x0;
  const float* left;
  const float* right;
  float* grad;
  float* x0_out;
  float* out2;
  const float* aux1;
  const float* aux2;
  int32_t B, C, H, W, margin;
  float disp_scale, disp_offset, min_disp, max_disp, w_ssim;
  float step, ca, cb, cc;
} d3_guidance_params;
int d3_guidance_step(const d3_guidance_params* p, void* stream);
/* grad[b,0,y,x] = d(loss)/d(disp_norm) with disparity = clip(min_disp + (disp_norm - shift) * (max_disp - min_disp)):
 * the output of the reference's FlowGuidance.grad.  disp_norm fp32 [B,1,H,W]. */
int d3_guidance_grad(const float* disp_norm, const float* left, const float* right, int B, int H, int W, int margin,
                     float min_disp, float max_disp, float shift, float w_ssim, float* grad, void* stream);

/* ---- pre/post-processing (fp32, elementwise) -------------------------------------------------
 * inference.py:128-131 / 246-249: out = fxb / in where in > 0 else 0. */
int d3_depth_to_disp(const float* depth, float fxb, int64_t n, float* disp, void* stream);
int d3_disp_to_depth(const float* disp, float fxb, int64_t n, float* depth, void* stream);
/* Normalizer (ssi=False, mode="average"; utils_d3roma/utils.py:183-196): disp [N][HW] -> out [N][num_chs][HW],
 * ((min(x, bound) / bound) ** gamma - t) * s replicated over num_chs. */
int d3_normalize_disp(const float* disp, int64_t N, int64_t HW, int num_chs, float bound, float gamma, float t,
                      float s, float* out, void* stream);
/* Normalizer.denormalize (utils.py:252-273): y [NR][num_chs][HW] -> out [NR][HW],
 * sum_c (y_c / s + t) ** (1/gamma) * (bound / num_chs). */
int d3_denormalize_disp(const float* y, int64_t NR, int64_t HW, int num_chs, float bound, float gamma, float t,
                        float s, float* out, void* stream);
/* Normalizer mode="piecewise" (utils_d3roma/utils.py:198-229, 275-293; the LDM-family configs): the disparity is split
 * over num_chs <= 3 channels, channel c covering bounds[c] pixels with its own gamma.  bounds / gammas are HOST arrays of
 * num_chs floats.  disp [N][HW] -> out [N][num_chs][HW];  y [NR][num_chs][HW] -> out [NR][HW]. */
int d3_normalize_piecewise(const float* disp, int64_t N, int64_t HW, int num_chs, const float* bounds,
                           const float* gammas, float t, float s, float* out, void* stream);
int d3_denormalize_piecewise(const float* y, int64_t NR, int64_t HW, int num_chs, const float* bounds,
                             const float* gammas, float t, float s, float* out, void* stream);
/* InputPadder.pad (replicate) / unpad (utils.py:313-325) over `planes` = B*C planes of H x W. */
int d3_pad_replicate(const float* in, int64_t planes, int H, int W, int left, int right, int top, int bottom,
                     float* out, void* stream);
int d3_crop(const float* in, int64_t planes, int Hin, int Win, int top, int left, int H, int W, float* out,
            void* stream);
/* InputPadder.pad_zero (utils.py:317-319; the training mask, train.py:186). */
int d3_pad_zero(const float* in, int64_t planes, int H, int W, int left, int right, int top, int bottom, float* out,
                void* stream);
/* train.py:175-178, in place: x[B][C][HW] <- clamp(x, -range, range); mask[B][1][HW] <- mask * (|x| < range). */
int d3_clamp_target(float* x, float* mask, int64_t B, int C, int64_t HW, float range, void* stream);
/* out[b][out_off + i] = clamp(in[b][i], lo, hi), i < chw (custom_pipelines.py:176-182,196: image.clamp(-1,1) written
 * straight into its slot of the intermediate-image stack). */
int d3_clamp_into(const float* in, int64_t B, int64_t chw, int64_t out_bstride, int64_t out_off, float lo, float hi,
                  float* out, void* stream);

/* ---- evaluation post-processing on the device ----------------------------------------------
 * compute_errors(space="disp") of utils_d3roma/utils.py:554-653: per sample, over mask != 0 pixels,
 * out[b] = {EPE, D1, D2, D3, D5 (thresholds scaled by n_valid/480 as the reference does), a1(1.05), a2(1.10), a3(1.25),
 * RMSE, REL, MAE of depth = fxb / disp}. fxb: fxb_n = 0 (1/disp), 1 (shared) or B (per sample) entries. */
int d3_disp_metrics(const float* gt_disp, const float* pred_disp, const float* mask, const float* fxb, int fxb_n, int B,
                    int64_t HW, double* out, void* stream);
/* compute_scale_and_shift (utils.py:488-528): per row least squares pred*s + t ~ target over mask; out [rows][2]. */
int d3_scale_shift_lsq(const float* pred, const float* target, const float* mask, int rows, int64_t HW,
                       float* scale_shift, void* stream);
/* Normalizer ssi=True (utils.py:171-182): per-row min/max over the mask, then (clamp((x-low)/(up-low),0,1)-t)*s on the
 * mask and 0 elsewhere; and the inverse affine `y*s + t` per row (utils.py:250-251). mask may be NULL (all valid). */
int d3_masked_minmax(const float* x, const float* mask, int rows, int64_t HW, float* low_up, void* stream);
int d3_normalize_ssi(const float* x, const float* mask, const float* low_up, int rows, int64_t HW, float t, float s,
                     float* out, void* stream);
int d3_affine_rows(const float* y, const float* scale_shift, int rows, int64_t HW, float* out, void* stream);

/* ---- one RANSAC iteration of the robust scale/shift fit (utils_d3roma/ransac.py:51-96, used by
 * Normalizer.denormalize with ssi + safe_ssi, utils.py:240-243): fit on the n_sample pixels perm[0..n_sample) (the
 * caller's random permutation of 0..HW-1, int32), inliers = ((target - (pred*s+t))^2 < T) & mask, refit on them, score
 * 1 / mean(max(Y/X'', X''/Y) < 1.25) over ALL rows and pixels, and per row keep the refit when it has more than
 * min_inliers inliers and a better score.  best_fit [rows][2] (initialise to {1, 0}), best_err [rows] (initialise to
 * +inf), best_num [rows] (0) are updated in place; workspace: d3_ransac_workspace_bytes(rows, HW) bytes of scratch. */
size_t d3_ransac_workspace_bytes(int rows, int64_t HW);
int d3_ransac_iteration(const float* pred, const float* target, const float* mask, const int* perm, int n_sample, int rows,
                        int64_t HW, float T, float min_inliers, void* workspace, float* best_fit, float* best_err,
                        float* best_num, void* stream);

/* ---- camera-frame front end (inference.py:99-213: cv2.resize + .float() + (x/255-0.5)*2 + batchify, padded as
 * InputPadder would).  frames: uint8 [B,Hs,Ws,Cimg] (Cimg 1 = gray tiled to 3, 3 = RGB, 4 = RGBA without alpha);
 * out: fp32 NCHW [B,3,Hp,Wp] where the H x W resized image sits at (pad_top, pad_left) and the border is replicated.
 * The bilinear resize is OpenCV's 8-bit fixed-point INTER_LINEAR, bit for bit. */
int d3_frame_to_input(const uint8_t* frames, int B, int Hs, int Ws, int Cimg, int H, int W, int pad_top, int pad_left,
                      int Hp, int Wp, float* out, void* stream);

/* ---- raw sensor depth (metres, fp32 [B,Hs,Ws]) -> cv2.resize(INTER_NEAREST) to H x W -> raw disparity
 * fxb / depth where depth > 0 else 0 (inference.py:184-186) [B,1,H,W], validity mask (raw_disp > 0) [B,1,H,W], and
 * the Normalizer("average")-normalised, replicate-padded conditioning map [B,1,Hp,Wp] (utils.py:171-196, 297-325). */
int d3_depth_to_input(const float* depth, int B, int Hs, int Ws, int H, int W, int pad_top, int pad_left, int Hp,
                      int Wp, float fxb, float bound, float gamma, float t, float s, float* raw_disp, float* mask,
                      float* norm_padded, void* stream);

/* ==== latent family: diffusers UNet2DConditionModel in the SD-2.1 configuration
 * (/root/reference/train.py:743-772, called at core/custom_pipelines.py:386) and AutoencoderKL (train.py:830-833, used at
 * custom_pipelines.py:35-59, 255-272).  Convolutions and every Linear layer run on d3_conv2d_igemm; the entry points below
 * are the remaining ops.  16-bit NHWC activations / [tokens][C] rows with explicit pitches, as above. ================== */

/* GroupNorm for any group width (aten::native_group_norm of ResnetBlock2D.norm1/2, Transformer2DModel.norm,
 * conv_norm_out): pass 1 accumulates per-channel sum / sum of squares of cat(x1, x2) into chan_stats fp64 [B][C1+C2][2]
 * (zero on entry); pass 2 normalises (+ SiLU) into one dense tensor. */
int d3_gn_stats(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2, int B, int HW,
                double* chan_stats, int dtype, void* stream);
int d3_gn_apply(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2, const double* chan_stats,
                const float* gamma, const float* beta, int B, int HW, int groups, float eps, int apply_silu, void* out,
                int out_pitch, int dtype, void* stream);
/* The same GroupNorm (+ SiLU) as ONE kernel, for tensors small enough to stay in the 126 MB L2 between its two passes (the
 * latent UNet's): a block owns one sample and a set of whole groups, reduces their moments itself and normalises - no
 * statistics buffer, no atomics.  Channels per group must be even. */
int d3_groupnorm_one_pass(const void* x1, int C1, int pitch1, const void* x2, int C2, int pitch2, const float* gamma,
                          const float* beta, int B, int HW, int groups, float eps, int apply_silu, void* out,
                          int out_pitch, int dtype, void* stream);
/* aten::native_layer_norm over the last dimension of M rows of C (<= 1280) channels (BasicTransformerBlock.norm1-3). */
int d3_layernorm(const void* x, int x_pitch, const float* gamma, const float* beta, int64_t M, int C, float eps,
                 void* out, int out_pitch, int dtype, void* stream);
/* diffusers GEGLU after its projection: out[m][d] = x[m][d] * gelu(x[m][D + d]), exact (erf) GELU. */
int d3_geglu(const void* x, int x_pitch, int64_t M, int D, void* out, int out_pitch, int dtype, void* stream);
/* softmax(q k^T / sqrt(64)) v, head_dim 64 (diffusers Attention of attn1): q / k / v rows of B*N tokens with their own
 * pitches (channel slices of a fused QKV tensor), head h at channels [64h, 64h + 64); out [B*N][out_pitch]. */
int d3_attention_hd64(const void* q, int q_pitch, const void* k, int k_pitch, const void* v, int v_pitch, int B, int N,
                      int heads, void* out, int out_pitch, int dtype, void* stream);
/* The same attention (head_dim 64) on the Blackwell tensor path: S = Q K^T and O = P V as tcgen05.mma with TMEM
 * accumulators, Q / K / V^T tiles by TMA, softmax rows read straight from TMEM.  `workspace` (device scratch of
 * d3_attention_hd64_tc_workspace_bytes(B, N, heads) bytes) receives V^T with the keys contiguous. */
size_t d3_attention_hd64_tc_workspace_bytes(int B, int N, int heads);
int d3_attention_hd64_tc(const void* q, int q_pitch, const void* k, int k_pitch, const void* v, int v_pitch, int B, int N,
                         int heads, void* out, int out_pitch, void* workspace, int dtype, void* stream);
/* attention of N query tokens per sample against L <= 8 context tokens whose keys / values kc, vc fp32 [B][L][C] were
 * projected beforehand (attn2 against the [B, 2, 1024] empty-prompt embedding, custom_pipelines.py:299-304, 386);
 * C = 64 * heads. */
int d3_cross_attention_small(const void* q, int q_pitch, const float* kc, const float* vc, int B, int N, int L, int C,
                             void* out, int out_pitch, int dtype, void* stream);
/* out[n][r] = act_out(bias[r] + sum_k w[r][k] * act_in(x[n][k])), fp32, act 0 = identity / 1 = SiLU / 2 = GELU (erf) /
 * 3 = quick GELU: the time-embedding MLP, all time_emb_proj layers stacked, the context key / value projections, the
 * Linear layers of the text encoder.  bias may be NULL; K % 4 == 0. */
int d3_linear_small(const float* x, int n_rows, int K, const float* w, const float* bias, int rows, int act_in,
                    int act_out, float* out, int out_stride, void* stream);
/* Output gather over NVLink peer memory (distributed_evaluate.py:162-163 `gather_for_metrics` of the predicted maps and
 * metric rows).  One process per GPU: d3_peer_alloc makes this rank's gather buffer (cudaMalloc, zeroed) and
 * its 64-byte CUDA IPC handle; the handles are exchanged by the host (torch.distributed) and d3_peer_open maps a peer's
 * buffer into this process (peer access over NVLink).  Per step:
 *   d3_denormalize_gather: Normalizer.denormalize (utils.py:252-273, as d3_denormalize_disp) with every value stored at
 *     element dst_elem_offset + i of EACH of the n_dst buffers (dst_bases: HOST array of device pointers, own buffer
 *     included) - the exchange is part of the last kernel of the loop;
 *   d3_peer_push: the same multi-destination store for final bytes (metric rows);
 *   d3_peer_signal: flag word `slot` of every peer's flag array := epoch (release, system scope), launched after the stores;
 *   d3_peer_wait: returns (in stream order) when the n local flag words are >= epoch; traps after ~20 s without a peer. */
int d3_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int d3_peer_open(const unsigned char* handle64, void** ptr);
int d3_peer_close(void* ptr);
int d3_peer_free(void* ptr);
int d3_denormalize_gather(const float* y, int64_t NR, int64_t HW, int num_chs, float bound, float gamma, float t, float s,
                          void* const* dst_bases, int n_dst, int64_t dst_elem_offset, void* stream);
int d3_peer_push(const void* src, size_t bytes, void* const* dst_bases, int n_dst, size_t dst_byte_offset, void* stream);
int d3_peer_signal(void* const* flag_bases, int n_dst, int slot, uint32_t epoch, void* stream);
int d3_peer_wait(const uint32_t* local_flags, int n, uint32_t epoch, void* stream);
/* ensemble_depths (utils_d3roma/utils.py:28-120; `config.ensemble`, evaluate.py:117-120): n <= 32 maps [n][P] of one
 * sample, per-map scale s and shift t (HOST arrays of n floats: the BFGS iterate of scipy lives on the host).
 *   d3_ensemble_objective: one evaluation of the closure of utils.py:64-87: out3 (device, 3 doubles) =
 *     { sum over pairs i < j and pixels of (v_i - v_j)^2,  min over pixels of pred,  max over pixels of pred },
 *     v_i = x_i * s_i + t_i, pred = lower median (reduction 1) or mean (reduction 0) over the maps; deterministic.
 *   d3_ensemble_reduce: utils.py:99-118: aligned [P] = median / mean of v, uncertainty [P] = median absolute deviation /
 *     unbiased std, both rescaled with the aligned map's min and max: (aligned - min) / (max - min), uncertainty / (max - min).
 * workspace: d3_ensemble_workspace_bytes() bytes of device scratch. */
size_t d3_ensemble_workspace_bytes(void);
int d3_ensemble_objective(const float* images, int n, int64_t P, const float* s_host, const float* t_host, int reduction,
                          void* workspace, double* out3, void* stream);
int d3_ensemble_reduce(const float* images, int n, int64_t P, const float* s_host, const float* t_host, int reduction,
                       void* workspace, float* aligned, float* uncertainty, void* stream);
/* Text encoder of the latent pipeline (custom_pipelines.py:260-287 `__encode_empty_text`: tokenizer("") ->
 * text_encoder(input_ids)[0]; transformers CLIPTextModel - embeddings, pre-LayerNorm transformer layers with a causal
 * mask, final LayerNorm), fp32 throughout; its Linear layers are d3_linear_small calls.
 *   d3_embed_tokens: out[n][:] = token_embedding[ids[n]][:] + position_embedding[n % seq_len][:]  (ids int64 on the device)
 *   d3_layernorm_f32: LayerNorm over the rows of fp32 [M][C]
 *   d3_causal_attention_f32: q / k / v fp32 rows of B*L tokens (row pitch `pitch`, head h at [h*head_dim, (h+1)*head_dim)):
 *     out[i] = softmax_j<=i(q_i . k_j * scale) v_j per head; L <= 128. */
int d3_embed_tokens(const int64_t* ids, int n, int seq_len, const float* token_embedding, int vocab,
                    const float* position_embedding, int max_positions, int D, float* out, void* stream);
int d3_layernorm_f32(const float* x, int64_t M, int C, const float* gamma, const float* beta, float eps, float* out,
                     void* stream);
int d3_causal_attention_f32(const float* q, const float* k, const float* v, int pitch, int B, int L, int heads,
                            int head_dim, float scale, float* out, int out_pitch, void* stream);
/* diffusers Timesteps(dim, flip_sin_to_cos, downscale_freq_shift): t fp32 [nt] -> out fp32 [nt][dim]. */
int d3_timestep_sincos(const float* t, int nt, int dim, int flip_sin_to_cos, float freq_shift, float* out, void* stream);
/* layout glue: fp32 NCHW [B,C,H,W] <-> channels [c_off, c_off + C) of a 16-bit NHWC tensor with pixel pitch `pitch`. */
int d3_nchw_to_nhwc16(const float* x, int B, int C, int H, int W, void* out, int out_pitch, int c_off, int dtype,
                      void* stream);
int d3_nhwc16_to_nchw(const void* x, int pitch, int c_off, int B, int C, int H, int W, float* out, int dtype,
                      void* stream);
/* F.interpolate(size=(Ho, Wo), mode="nearest") of a 16-bit NHWC tensor into the top-left Ho x Wo window of a
 * [B, Hbuf, Wbuf, C] buffer; the rest of the buffer is zero-filled (Upsample2D with an explicit output size, and the
 * even-sized zero-extended copy an odd feature map needs before a stride-2 conv). */
int d3_resize_nearest_nhwc16(const void* in, int B, int H, int W, int C, int in_pitch, void* out, int Ho, int Wo,
                             int Hbuf, int Wbuf, int out_pitch, void* stream);
/* VAE mid-block attention (diffusers Attention with one 512-wide head, AutoencoderKL mid_block.attentions.0) runs as two
 * d3_conv2d_igemm GEMMs per sample - S = Q K^T with K as the "weight" operand, O = P V with V^T - around these helpers:
 * in-place row softmax of scale * S over the first n_cols columns of a 16-bit [rows][pitch] matrix (padding columns
 * are zeroed), and V [N][C] -> V^T [C][Np] zero-padded to the K-block multiple. */
int d3_softmax_rows(void* s, int rows, int n_cols, int pitch, float scale, int dtype, void* stream);
int d3_transpose_pad16(const void* in, int N, int C, int in_pitch, void* out, int Np, void* stream);
/* quant_conv (8 -> 8) / post_quant_conv (4 -> 4) of AutoencoderKL (custom_pipelines.py:54, 268): 1x1 conv over fp32 NCHW
 * with at most 8 channels on either side; w fp32 [Cout][Cin]. */
int d3_conv1x1_nchw_small(const float* x, int B, int Cin, int Cout, int64_t HW, const float* w, const float* bias,
                          float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step: backward pass, loss and optimizer of /root/reference/train.py:157-538
 * (`accelerator.backward(loss)` :455, `clip_grad_norm_` :518, `optimizer.step()` :524 over torch.optim.AdamW :862).
 * The data gradient of a convolution is a convolution with transposed weights and runs on d3_conv2d_igemm (weights
 * re-packed by d3_pack_taps16); the entry points below are what autograd / the optimizer add to the forward kernels.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Weight gradient of a 3x3 (padding 1, stride 1 or 2) or 1x1 convolution - replaces the weight half of
 * aten::convolution_backward:   dw[co][col0 + tap*Cin + ci] += sum_{b,y,x} dy[b,y,x,co] * x[b, s*y+kh-1, s*x+kw-1, ci]
 * (the [Cout][taps*Cin] layout of the forward weight packs, in fp32, ACCUMULATED: zero it first). */
typedef struct d3_wgrad_params {
  const void* x;       /* 16-bit NHWC [B,H,W,Cin]: the conv's input, pixel pitch x_pitch                  */
  int32_t B, H, W;     /* INPUT spatial size                                                               */
  int32_t Cin;         /* multiple of 64                                                                   */
  int32_t x_pitch;
  const void* dy;      /* 16-bit NHWC [B,H/stride,W/stride,Cout]: gradient w.r.t. the conv's output        */
  int32_t Cout;        /* multiple of 128                                                                  */
  int32_t dy_pitch;
  int32_t ksize;       /* 3 or 1                                                                           */
  int32_t stride;      /* 1, or 2 (ksize 3 only)                                                           */
  float* dw;           /* fp32 [Cout][dw_pitch]                                                            */
  int64_t dw_pitch;    /* row pitch of dw in elements (multiple of 4)                                      */
  int32_t dw_col0;     /* first column of this conv's block inside the row (multiple of 4)                 */
  int32_t dtype;       /* D3_DTYPE_BF16 / D3_DTYPE_F16                                                     */
} d3_wgrad_params;
int d3_conv2d_wgrad(const d3_wgrad_params* p, void* stream);

/* Backward of d3_groupnorm_silu (native_group_norm_backward + silu_backward) over the virtual concat of two sources:
 * dy = gradient w.r.t. the dense output; dx (dense, C1 + C2 channels) = gradient w.r.t. the concatenated input
 * (+ up to 3 `addends`: dense 16-bit tensors of the same shape with their own pixel pitches - the gradients that reached
 * the input over other paths, e.g. the residual branch); dgamma / dbeta fp32 [C] are ACCUMULATED; `colsum` (optional, fp32
 * rows `colsum_bstride` elements apart, accumulated; stride 0 = all samples into one row) receives the per-(sample,
 * channel) sums of dx over the first `colsum_channels` channels (0 = all) - the bias / time-embedding gradient of the
 * conv that produced the input.  workspace: d3_groupnorm_bwd_workspace_bytes(B, C1 + C2) bytes. */
size_t d3_groupnorm_bwd_workspace_bytes(int B, int C);
int d3_groupnorm_silu_bwd(const void* x1, int C1, int p1, const double* stats1, const void* x2, int C2, int p2,
                          const double* stats2, const float* gamma, const float* beta, int B, int HW, int groups,
                          float eps, int apply_silu, const void* dy, int pdy, const void* const* addends,
                          const int* addend_pitches, int n_addends, void* dx, int pdx, float* dgamma, float* dbeta,
                          float* colsum, int64_t colsum_bstride, int colsum_channels, float* workspace, int dtype,
                          void* stream);
/* out[b * out_bstride + c] += sum over pixels of x[b,p,c] (bias gradients; out_bstride 0 sums over the batch too); out = a + b; backward of the nearest-2x upsample
 * (2x2 block sums of a [B,2H,2W,C] tensor, + optional addend). */
int d3_colsum_nhwc16(const void* x, int pitch, int B, int HW, int C, float* out, int64_t out_bstride, int dtype,
                     void* stream);
int d3_add_nhwc16(const void* a, int pa, const void* b, int pb, void* out, int po, int64_t npix, int C, int dtype,
                  void* stream);
/* out = a + b and colsum[b * colsum_bstride + c] += per-channel sums of the result, in one pass. */
int d3_add_colsum_nhwc16(const void* a, int pa, const void* b, int pb, void* out, int po, int B, int HW, int C,
                         float* colsum, int64_t colsum_bstride, int dtype, void* stream);
int d3_sumpool2x_nhwc16(const void* in, int pin, const void* addend, int padd, void* out, int pout, int B, int H, int W,
                        int C, int dtype, void* stream);
/* Backward of d3_conv_out_fused (conv_norm_out + SiLU + conv_out): dy fp32 NCHW [B,dc,H,W] -> da = gradient w.r.t.
 * SiLU(GroupNorm(h)) as 16-bit NHWC [B,H,W,128] (feed it to d3_groupnorm_silu_bwd), dw fp32 [dc][9][128] accumulated. */
int d3_conv_out_bwd(const void* h, int pitch, const double* stats, const float* gamma, const float* beta, const float* w,
                    const float* dy, int B, int H, int W, int dc, float eps, void* da, int pda, float* dw, int dtype,
                    void* stream);
/* Backward of d3_attention_hd8: dq | dk | dv in the layout of the QKV tensor.  workspace:
 * d3_attention_hd8_bwd_workspace_bytes(B, N, C) bytes. */
size_t d3_attention_hd8_bwd_workspace_bytes(int B, int N, int C);
int d3_attention_hd8_bwd(const void* qkv, int pitch, const void* o, int opitch, const void* d_o, int dopitch, int B, int N,
                         int C, void* dqkv, int dpitch, float* workspace, int dtype, void* stream);
/* The loss of train.py:322-417: network_loss = loss_fn(pred * mask, target * mask) (MSE, or L1 when l1 != 0; mask fp32
 * [B,1,H,W] or NULL), loss = mean(network_loss * weights[b]).  loss_sum[b] (fp64, accumulated) = sum of network_loss over
 * sample b; dy = grad_scale * weights[b] * d network_loss / d pred (grad_scale = loss_scale / (B*dc*plane));
 * dbias[d] (optional, accumulated) = sum of dy over channel d (conv_out's bias gradient). */
int d3_masked_loss_grad(const float* pred, const float* target, const float* mask, const float* weights, int B, int dc,
                        int64_t plane, int l1, float grad_scale, float* dy, double* loss_sum, float* dbias, void* stream);
/* Backward of d3_linear_small, y = act_in(x) W^T: dw [rows][K] and db [rows] accumulated (dw NULL: skip),
 * dx [n_rows][K] = (dy W) * act_in'(x) (dx NULL: skip; accumulate_dx: add to dx). */
int d3_linear_small_bwd(const float* x, int n_rows, int K, int act_in, const float* w, int rows, const float* dy,
                        float* dw, float* db, float* dx, int accumulate_dx, void* stream);
/* pyramid_noise_like (utils_d3roma/utils.py:122-133; `noise_strategy: pyramid` of the pixel-space training configs):
 * dst[planes][H][W] += mul * nn.Upsample(size=(H, W), mode="bilinear")(src[planes][h][w]);  x /= x.std() (unbiased;
 * workspace2: two doubles). */
int d3_bilinear_upsample_acc(const float* src, int64_t planes, int h, int w, float* dst, int H, int W, float mul,
                             void* stream);
int d3_div_by_std(float* x, int64_t n, double* workspace2, void* stream);
/* Optimizer (torch.optim.AdamW semantics, train.py:862; clip_grad_norm_ :518): out += sum g^2 (fp64);  one fused AdamW
 * step over a flat parameter buffer - gradients are multiplied by inv_loss_scale and by min(1, max_norm / (norm + 1e-6))
 * with norm = sqrt(*sumsq) * inv_loss_scale (sumsq NULL or max_norm <= 0: no clipping); the first n16 parameters are
 * also written as 16-bit (the tensor-core weight packs); norm_out (optional) receives the gradient norm.  A non-finite norm
 * (sumsq given) leaves parameters and moments untouched - the step is skipped, as torch's GradScaler does. */
int d3_sumsq(const float* g, int64_t n, double* out, void* stream);
int d3_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, void* p16, int64_t n16, int dtype16, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, const double* sumsq,
                  float inv_loss_scale, float max_norm, float* norm_out, void* stream);
int d3_cast_f32_to_16(const float* x, void* out, int64_t n, int dtype16, void* stream);
/* out[i] = src[ia[i]] + (ib[i] >= 0 ? src[ib[i]] : 0): the merged per-channel biases the forward kernels read. */
int d3_gather_add(const float* src, const int32_t* ia, const int32_t* ib, float* out, int n, void* stream);
/* Data-gradient weight packs: out[ci][j][co] (16-bit) = tap_map[j] >= 0 ? w[co][col0 + tap_map[j]*Cin + ci] : 0
 * (flipped taps for a stride-1 conv, the phase sub-sets of a stride-2 conv, the plain transpose for 1x1). */
int d3_pack_taps16(const float* w, int64_t w_pitch, int col0, int Cin, int Cout, const int32_t* tap_map, int n_out_taps,
                   void* out, int dtype16, void* stream);
/* The same for every pack of a model in one launch: `entries_dev` / `tile_prefix_dev` are DEVICE arrays (n_entries
 * entries; tile_prefix[i] = number of 32 x 32 tiles, Cout/32 * Cin/32 * n_out_taps each, of the entries before i;
 * total_tiles = their sum); offsets are in elements from w_base (fp32 master buffer) / out_base (16-bit pack buffer). */
typedef struct d3_pack_entry {
  int64_t w_offset;
  int64_t w_pitch;
  int64_t out_offset;
  int32_t col0, Cin, Cout, n_out_taps;
  int32_t tap_map[16];
} d3_pack_entry;
int d3_pack_taps16_batched(const d3_pack_entry* entries_dev, const int32_t* tile_prefix_dev, int n_entries, int total_tiles,
                           const float* w_base, void* out_base, int dtype16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D3ROMA_B200_H */
