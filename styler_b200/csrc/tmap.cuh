// Host-side TMA descriptor (CUtensorMap) construction with a small cache.  The driver entry point is
// resolved at run time through the CUDA runtime, so the library has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// Tiled tensor map over a `rank`-D tensor (dims/strides innermost first; strides in BYTES for dims 1..rank-1),
// 128-byte swizzle, out-of-bounds elements read as zero.  elem: 0 = fp32 (loaded as TF32-rounded), 1 = bf16,
// 2 = fp32 (plain).  Returns 0 or a negative error (message set).
int make_tmap(CUtensorMap* out, const void* base, int elem, int rank, const uint64_t* dims, const uint64_t* strides,
              const uint32_t* box);

}  // namespace sb
