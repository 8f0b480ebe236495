// Developer probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b.x4 and the smem image written by
// stmatrix.x4.trans from those fragments.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_tmem_layout.bin
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(uint32_t* out_regs, uint16_t* out_smem) {
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(1024) uint16_t tile[64 * 64];   // 64 pixel rows x 64 channels (128 B rows)
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_ptr;
  // every warp writes its 32-lane quarter: value = float(lane_abs * 256 + col)  (exact in bf16? no - keep fp32)
  {
    const uint32_t lane_abs = warp * 32 + lane;
    for (int c = 0; c < 64; c += 4) {
      uint32_t v0 = __float_as_uint((float)(lane_abs * 256 + c)), v1 = __float_as_uint((float)(lane_abs * 256 + c + 1));
      uint32_t v2 = __float_as_uint((float)(lane_abs * 256 + c + 2)), v3 = __float_as_uint((float)(lane_abs * 256 + c + 3));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((warp * 32) << 16) + c),
                   "r"(v0), "r"(v1), "r"(v2), "r"(v3));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) tile[i] = 0xFFFF;
  __syncthreads();
  if (warp == 1) {   // quarter 1: lanes 32..63; load lanes 32+16 .. 32+31 (second half), columns 0..31
    uint32_t r[16];
    const uint32_t taddr = base + ((warp * 32 + 16) << 16) + 0;
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 16; ++i) out_regs[lane * 16 + i] = r[i];
    // stmatrix.x4.trans of column blocks 0 and 1: matrices {lo(0), hi(0), lo(1), hi(1)}; value stored = lane index
    // (the TMEM lane = "channel") * 64 + column ("pixel") encoded so that it fits in 16 bits: (lane_abs - 48) * 64 + col
    uint32_t pk[4];
    for (int m = 0; m < 4; ++m) {
      const int i = m >> 1, hl = m & 1;
      const float a = __uint_as_float(r[4 * i + 2 * hl]), b = __uint_as_float(r[4 * i + 2 * hl + 1]);
      const uint32_t ea = (uint32_t)(((int)a / 256 - 48) * 64 + ((int)a % 256));
      const uint32_t eb = (uint32_t)(((int)b / 256 - 48) * 64 + ((int)b % 256));
      pk[m] = (ea & 0xFFFF) | (eb << 16);
    }
    const int m = lane >> 3, j = lane & 7;
    const int hl = m & 1, ii = m >> 1;
    // row = pixel 8*ii + j, channels chunk hl (8 channels = 16 B), no swizzle in the probe
    const uint32_t addr = smem_u32(tile) + (8 * ii + j) * 128 + hl * 16;
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]),
                 "r"(pk[1]), "r"(pk[2]), "r"(pk[3]));
    // read back through ldmatrix.x4.trans with the same addresses: must reproduce pk[]
    uint32_t rb[4];
    __syncwarp();
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]) : "r"(addr));
    int bad = 0;
    for (int k = 0; k < 4; ++k) bad |= (rb[k] != pk[k]);
    if (bad) printf("ldmatrix.trans mismatch lane %d: %08x %08x %08x %08x vs %08x %08x %08x %08x\n", lane, rb[0], rb[1],
                    rb[2], rb[3], pk[0], pk[1], pk[2], pk[3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out_smem[i] = tile[i];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64));
}

int main() {
  uint32_t* d_regs; uint16_t* d_smem;
  cudaMalloc(&d_regs, 32 * 16 * 4); cudaMalloc(&d_smem, 64 * 64 * 2);
  probe<<<1, 128>>>(d_regs, d_smem);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  uint32_t regs[32 * 16]; uint16_t sm[64 * 64];
  cudaMemcpy(regs, d_regs, sizeof(regs), cudaMemcpyDeviceToHost);
  cudaMemcpy(sm, d_smem, sizeof(sm), cudaMemcpyDeviceToHost);
  int ok = 1;
  for (int t = 0; t < 32; ++t) {
    for (int i = 0; i < 16; ++i) {
      float f; memcpy(&f, &regs[t * 16 + i], 4);
      const int lane = (int)f / 256, col = (int)f % 256;
      // expected: rep = i/4; lane = 48 + t/4 + 8*((i%4)/2); col = 8*rep + 2*(t%4) + (i%2)
      const int el = 48 + t / 4 + 8 * ((i % 4) / 2), ec = 8 * (i / 4) + 2 * (t % 4) + (i % 2);
      if (lane != el || col != ec) { ok = 0; if (t < 6) printf("t%d r%d: lane %d col %d (expected %d %d)\n", t, i, lane, col, el, ec); }
    }
  }
  printf("tcgen05.ld.16x256b.x4 layout as expected: %s\n", ok ? "YES" : "NO");
  // smem image: row = pixel (0..15), 16-bit element index e (0..15) = channel-in-16; expect value = ch * 64 + pixel
  int ok2 = 1;
  for (int p = 0; p < 16; ++p)
    for (int c = 0; c < 16; ++c) {
      const int v = sm[p * 64 + c];
      if (v != c * 64 + p) { ok2 = 0; if (p < 3) printf("smem[pixel %d][ch %d] = %d (expected %d)\n", p, c, v, c * 64 + p); }
    }
  printf("stmatrix.x4.trans image [pixel][channel] as expected: %s\n", ok2 ? "YES" : "NO");
  return 0;
}
