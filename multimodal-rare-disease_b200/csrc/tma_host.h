// Host-side construction of TMA tensor maps.  cuTensorMapEncodeTiled is resolved through the
// runtime's driver-entry-point query so the library has no link-time dependency on libcuda.so
// (it must load on machines without a driver; only compute calls need the GPU).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mrd {

// rank in [1,5]; dims[rank] in elements (innermost first); strides_bytes[rank-1] for dims 1..rank-1;
// box[rank] in elements.  swizzle_bytes in {0, 32, 64, 128}.  elem_bytes in {2 (bf16), 4 (f32)}.
// Returns 0 on success, otherwise a CUresult / cudaError code (message via mrd_last_error()).
int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      int swizzle_bytes);

void set_last_error(const char* fmt, ...);
const char* get_last_error();

}  // namespace mrd
