// Host-side helpers to build TMA tensor maps without linking libcuda directly
// (the driver entry point is resolved at run time through the CUDA runtime, so the
// shared library still loads on a machine that has no GPU driver installed).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace d3 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled();

// Generic 16-bit (bf16 / fp16) tiled map with 128B swizzle and zero OOB fill.
//   rank      : 2..5
//   dims      : extent of each dimension, fastest first (elements)
//   strides_b : byte stride of dimensions 1..rank-1 (dimension 0 is contiguous)
//   box       : box extent per dimension (elements); box[0]*2 must be <= 128 bytes
// Returns 0 on success.
int make_tmap_16bit(bool f16, CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                   const uint32_t* box, bool swizzle128 = true);

}  // namespace d3
