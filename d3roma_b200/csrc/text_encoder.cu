// CLIP text encoder of the latent pipeline (custom_pipelines.py:260-287: `__encode_empty_text` tokenises "" and runs
// `text_encoder(input_ids)[0]`): the pieces of transformers' CLIPTextModel that the small fp32 Linear kernel
// (d3_linear_small, ldm_ops.cu) does not already cover.  The prompt is a handful of tokens (two for the empty prompt:
// BOS, EOS), the result is a constant of the checkpoint computed once per pipeline - so everything here is plain fp32
// CUDA-core work sized for <= 77 tokens, exact enough to be compared with the library at 1e-4.
//
//   d3_embed_tokens          out[n] = token_embedding[ids[n]] + position_embedding[n % seq_len]
//   d3_layernorm_f32         LayerNorm over the rows of an fp32 [M][C] matrix (two-pass, fp32 like torch's CPU kernel)
//   d3_causal_attention_f32  softmax(q k^T * scale + causal mask) v per head; one warp per (sequence, head, query)
#include "common.cuh"
#include "internal.h"

namespace d3 {

__global__ void __launch_bounds__(256) embed_tokens_kernel(const int64_t* __restrict__ ids, int n, int seq_len,
                                                           const float* __restrict__ tok, const float* __restrict__ pos,
                                                           int D, float* __restrict__ out) {
  const int row = blockIdx.x;
  const int64_t id = ids[row];
  const float* t = tok + (size_t)id * D;
  const float* p = pos + (size_t)(row % seq_len) * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[(size_t)row * D + i] = t[i] + p[i];
}

__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, int C,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float bc;
  const float* xr = x + (size_t)blockIdx.x * C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto block_sum = [&](float v) -> float {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      bc = t;
    }
    __syncthreads();
    const float r = bc;
    __syncthreads();
    return r;
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) s += xr[i];
  const float mean = block_sum(s) / (float)C;
  float q = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) {
    const float d = xr[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(block_sum(q) / (float)C + eps);
  for (int i = threadIdx.x; i < C; i += 256)
    out[(size_t)blockIdx.x * C + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}

// One warp per (sequence b, head h, query i); keys j <= i only (the causal mask of CLIPTextTransformer).  L <= 128.
__global__ void __launch_bounds__(128) causal_attention_f32_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                   const float* __restrict__ v, int pitch, int B, int L,
                                                                   int heads, int hd, float scale,
                                                                   float* __restrict__ out, int out_pitch) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * heads * L) return;
  const int i = warp % L, h = (warp / L) % heads, b = warp / (L * heads);
  const float* qr = q + ((size_t)b * L + i) * pitch + h * hd;
  float sc[4];   // scores of keys lane, lane + 32, lane + 64, lane + 96
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = lane + 32 * t;
    sc[t] = -INFINITY;
    if (j <= i) {
      const float* kr = k + ((size_t)b * L + j) * pitch + h * hd;
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(qr[d], kr[d], a);
      sc[t] = a * scale;
      mx = fmaxf(mx, sc[t]);
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    sc[t] = (lane + 32 * t <= i) ? expf(sc[t] - mx) : 0.f;
    sum += sc[t];
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float inv = 1.0f / sum;
  // out[d] = sum_j p_j v[j][d]: lanes stride over d, the probabilities come round by shuffle
  for (int d0 = 0; d0 < hd; d0 += 32) {
    const int d = d0 + lane;
    float acc = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float p = __shfl_sync(0xffffffffu, sc[j >> 5], j & 31);
      if (d < hd) acc = fmaf(p, v[((size_t)b * L + j) * pitch + h * hd + d], acc);
    }
    if (d < hd) out[((size_t)b * L + i) * out_pitch + h * hd + d] = acc * inv;
  }
}

}  // namespace d3

using namespace d3;

extern "C" int d3_embed_tokens(const int64_t* ids, int n, int seq_len, const float* token_embedding, int vocab,
                               const float* position_embedding, int max_positions, int D, float* out, void* stream) {
  D3_REQUIRE(ids && token_embedding && position_embedding && out, "embed_tokens: null pointer");
  D3_REQUIRE(n > 0 && seq_len > 0 && n % seq_len == 0 && D > 0 && vocab > 0, "embed_tokens: bad sizes");
  D3_REQUIRE(seq_len <= max_positions, "embed_tokens: %d tokens exceed the %d positions of the model", seq_len,
             max_positions);
  embed_tokens_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, n, seq_len, token_embedding,
                                                                        position_embedding, D, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_layernorm_f32(const float* x, int64_t M, int C, const float* gamma, const float* beta, float eps,
                                float* out, void* stream) {
  D3_REQUIRE(x && gamma && beta && out && M > 0 && M < (1 << 30) && C > 0, "layernorm_f32: bad arguments");
  layernorm_f32_kernel<<<(unsigned)M, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, C, gamma, beta, eps, out);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_causal_attention_f32(const float* q, const float* k, const float* v, int pitch, int B, int L, int heads,
                                       int head_dim, float scale, float* out, int out_pitch, void* stream) {
  D3_REQUIRE(q && k && v && out, "causal_attention_f32: null pointer");
  D3_REQUIRE(B > 0 && L > 0 && L <= 128 && heads > 0 && head_dim > 0, "causal_attention_f32: bad sizes (L <= 128)");
  D3_REQUIRE(pitch >= heads * head_dim && out_pitch >= heads * head_dim, "causal_attention_f32: bad pitch");
  const int warps = B * heads * L;
  causal_attention_f32_kernel<<<ceil_div(warps, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      q, k, v, pitch, B, L, heads, head_dim, scale, out, out_pitch);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
