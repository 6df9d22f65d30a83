// Host-side helper: build TMA tensor maps through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <stdint.h>
#include "../../include/opental_b200.h"

namespace otal {
// bf16 tensor, `rank` dims innermost-first; strides[i] = byte stride of dim i+1; box = tile extents.
// swizzle128: 0 = no swizzle, 1 = SWIZZLE_128B, 2 = SWIZZLE_64B.
int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);
int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swz);
}  // namespace otal
