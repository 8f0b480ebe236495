// Output gather over NVLink peer memory ("one all-gather per batch of the final disparity ... so the last
// kernel of the loop writes its slot directly"; distributed_evaluate.py:162-163 `gather_for_metrics`).
//
// One process per GPU.  Every rank owns a gather buffer allocated here with cudaMalloc and exported with a CUDA IPC
// handle; the handles travel through torch.distributed once (host side, d3roma_b200/distributed.py::PeerGather), every
// rank opens its peers' buffers (cudaIpcOpenMemHandle enables peer access over NVLink / NVSwitch) and from then on the
// exchange step is part of the LAST KERNEL of the denoising loop:
//
//   d3_denormalize_gather   Normalizer.denormalize (utils.py:252-273) of this rank's maps, each value computed once and
//                           stored into the rank's slot of EVERY rank's buffer (1 local + n-1 NVLink stores, coalesced)
//   d3_peer_push            the same for bytes that are already final (the [n, 11] metric rows)
//   d3_peer_signal          one release-store per peer: "slot r of epoch e is complete" (launched after the stores, so the
//                           kernel boundary orders it behind them)
//   d3_peer_wait            one thread per peer spins (acquire loads, bounded) until all slots of the epoch have arrived;
//                           kernels behind it on the stream read the gathered buffer
//
// No NCCL call, no staging copy, no host synchronisation in the step.  Buffers are double-buffered by epoch parity: a
// rank pushes epoch e + 1 only after it has consumed epoch e - 1's buffer in stream order, and nobody can be pushing
// epoch e + 1 before everybody pushed e.
#include <cstring>

#include "common.cuh"
#include "internal.h"

namespace d3 {

constexpr int kMaxPeers = 16;

#define D3_GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

struct PeerPtrs {
  void* p[kMaxPeers];
  int n;
};

__global__ void __launch_bounds__(256) denormalize_gather_kernel(const float* __restrict__ y, int64_t NR, int64_t HW,
                                                                int num_chs, float bound, float gamma, float t, float s,
                                                                const PeerPtrs dst, int64_t dst_off) {
  const float per = bound / (float)num_chs;
  const float inv_gamma = 1.0f / gamma;
  D3_GRID_STRIDE(i, NR * HW) {
    const int64_t n = i / HW, p = i - n * HW;
    float z = 0.f;
    for (int c = 0; c < num_chs; ++c) {     // the arithmetic of denormalize_disp_kernel (prepost.cu), rounding for rounding
      float v = __fadd_rn(__fdiv_rn(y[(n * num_chs + c) * HW + p], s), t);
      if (gamma != 1.0f) v = powf(v, inv_gamma);
      z = __fadd_rn(z, __fmul_rn(v, per));
    }
#pragma unroll 4
    for (int k = 0; k < dst.n; ++k) static_cast<float*>(dst.p[k])[dst_off + i] = z;
  }
}

// num_chs == 1: elementwise, four values per thread, 16-byte loads and (local + NVLink) stores
__global__ void __launch_bounds__(256) denormalize_gather_vec4_kernel(const float4* __restrict__ y, int64_t n4, float bound,
                                                                     float gamma, float t, float s, const PeerPtrs dst,
                                                                     int64_t dst_off4) {
  const float inv_gamma = 1.0f / gamma;
  auto f = [&](float x) {
    float v = __fadd_rn(__fdiv_rn(x, s), t);
    if (gamma != 1.0f) v = powf(v, inv_gamma);
    return __fadd_rn(0.f, __fmul_rn(v, bound));
  };
  D3_GRID_STRIDE(i, n4) {
    const float4 a = __ldcs(y + i);
    const float4 z = make_float4(f(a.x), f(a.y), f(a.z), f(a.w));
#pragma unroll 4
    for (int k = 0; k < dst.n; ++k) static_cast<float4*>(dst.p[k])[dst_off4 + i] = z;
  }
}

// 16-byte vectors when everything is aligned, bytes otherwise
__global__ void __launch_bounds__(256) peer_push_kernel(const unsigned char* __restrict__ src, size_t bytes,
                                                       const PeerPtrs dst, size_t dst_off, int vec) {
  if (vec) {
    const size_t nv = bytes >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
      const uint4 v = s4[i];
      for (int k = 0; k < dst.n; ++k) reinterpret_cast<uint4*>(static_cast<unsigned char*>(dst.p[k]) + dst_off)[i] = v;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x) {
      const unsigned char v = src[i];
      for (int k = 0; k < dst.n; ++k) static_cast<unsigned char*>(dst.p[k])[dst_off + i] = v;
    }
  }
}

__global__ void peer_signal_kernel(const PeerPtrs flags, int slot, unsigned epoch) {
  const int k = threadIdx.x;
  if (k >= flags.n) return;
  __threadfence_system();
  unsigned* f = static_cast<unsigned*>(flags.p[k]) + slot;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}

// Spins at most ~20 s (at 1 GHz) and then traps: a missing peer surfaces as a CUDA error, not as a hung GPU.
__global__ void peer_wait_kernel(const unsigned* __restrict__ local_flags, int n, unsigned epoch) {
  const int k = threadIdx.x;
  if (k < n) {
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_flags + k) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 20000000000LL) __trap();
      __nanosleep(200);
    } while (true);
  }
  __syncthreads();
  __threadfence_system();
}

static int fill_ptrs(PeerPtrs& pp, void* const* bases, int n, const char* what) {
  if (!(bases && n >= 1 && n <= kMaxPeers)) return set_error(D3_ERR_INVALID, "%s: 1 <= peers <= %d expected (got %d)", what, kMaxPeers, n);
  pp.n = n;
  for (int k = 0; k < n; ++k) {
    if (!bases[k]) return set_error(D3_ERR_INVALID, "%s: null peer pointer %d", what, k);
    pp.p[k] = bases[k];
  }
  return D3_OK;
}

}  // namespace d3

using namespace d3;

extern "C" int d3_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  D3_REQUIRE(bytes > 0 && ptr && handle64, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  D3_CHECK_CUDA(cudaMalloc(ptr, bytes));
  D3_CHECK_CUDA(cudaMemset(*ptr, 0, bytes));
  D3_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  D3_CHECK_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return D3_OK;
}

extern "C" int d3_peer_open(const unsigned char* handle64, void** ptr) {
  D3_REQUIRE(handle64 && ptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  D3_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return D3_OK;
}

extern "C" int d3_peer_close(void* ptr) {
  D3_REQUIRE(ptr, "peer_close: null pointer");
  D3_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return D3_OK;
}

extern "C" int d3_peer_free(void* ptr) {
  D3_REQUIRE(ptr, "peer_free: null pointer");
  D3_CHECK_CUDA(cudaFree(ptr));
  return D3_OK;
}

extern "C" int d3_denormalize_gather(const float* y, int64_t NR, int64_t HW, int num_chs, float bound, float gamma,
                                     float t, float s, void* const* dst_bases, int n_dst, int64_t dst_elem_offset,
                                     void* stream) {
  D3_REQUIRE(y && num_chs >= 1 && bound > 0.f && gamma > 0.f && s != 0.f && dst_elem_offset >= 0,
             "denormalize_gather: bad arguments");
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, dst_bases, n_dst, "denormalize_gather")) return rc;
  if (NR * HW == 0) return D3_OK;
  bool vec = num_chs == 1 && (NR * HW) % 4 == 0 && dst_elem_offset % 4 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0;
  for (int k = 0; k < n_dst; ++k) vec = vec && reinterpret_cast<uintptr_t>(dst_bases[k]) % 16 == 0;
  if (vec) {
    const int64_t n4 = NR * HW / 4, w4 = (n4 + 255) / 256;
    const int b4 = (int)(w4 < (int64_t)num_sms() * 16 ? w4 : (int64_t)num_sms() * 16);
    denormalize_gather_vec4_kernel<<<b4, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(y), n4, bound, gamma, t, s, pp, dst_elem_offset / 4);
    D3_LAUNCH_CHECK();
    return D3_OK;
  }
  const int64_t want = (NR * HW + 255) / 256;
  const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  denormalize_gather_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, NR, HW, num_chs, bound, gamma, t, s,
                                                                                  pp, dst_elem_offset);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_peer_push(const void* src, size_t bytes, void* const* dst_bases, int n_dst, size_t dst_byte_offset,
                            void* stream) {
  D3_REQUIRE(src, "peer_push: null source");
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, dst_bases, n_dst, "peer_push")) return rc;
  if (bytes == 0) return D3_OK;
  int vec = (bytes % 16 == 0) && (dst_byte_offset % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  for (int k = 0; k < n_dst; ++k) vec = vec && (reinterpret_cast<uintptr_t>(dst_bases[k]) % 16 == 0);
  const size_t units = vec ? bytes / 16 : bytes;
  const size_t want = (units + 255) / 256;
  const int blocks = (int)(want < (size_t)num_sms() * 8 ? want : (size_t)num_sms() * 8);
  peer_push_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char*>(src), bytes,
                                                                         pp, dst_byte_offset, vec);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_peer_signal(void* const* flag_bases, int n_dst, int slot, uint32_t epoch, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, flag_bases, n_dst, "peer_signal")) return rc;
  D3_REQUIRE(slot >= 0 && slot < kMaxPeers, "peer_signal: bad slot %d", slot);
  peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pp, slot, epoch);
  D3_LAUNCH_CHECK();
  return D3_OK;
}

extern "C" int d3_peer_wait(const uint32_t* local_flags, int n, uint32_t epoch, void* stream) {
  D3_REQUIRE(local_flags && n >= 1 && n <= kMaxPeers, "peer_wait: bad arguments");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(local_flags, n, epoch);
  D3_LAUNCH_CHECK();
  return D3_OK;
}
