#include "tma_host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

namespace mrd {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

const char* get_last_error() { return g_last_error; }

static PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?): %s",
                       cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    return fn;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      int swizzle_bytes) {
    auto fn = resolve_encode();
    if (!fn) return -2;
    CUtensorMapDataType dt =
        elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr,
                    bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error(
            "cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] "
            "strides=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u] swizzle=%d base=%p",
            (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
            (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
            (unsigned long long)(rank > 4 ? dims[4] : 0),
            (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
            (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
            (unsigned long long)(rank > 3 ? strides_bytes[2] : 0),
            (unsigned long long)(rank > 4 ? strides_bytes[3] : 0), box[0], rank > 1 ? box[1] : 0,
            rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, swizzle_bytes, base);
        return -1000 - static_cast<int>(r);
    }
    return 0;
}

}  // namespace mrd
