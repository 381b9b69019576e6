// Experiment: can a tcgen05 SWIZZLE_128B K-major A descriptor start at an arbitrary 128-byte row of a
// 1024-byte-aligned swizzled region (row-shifted views for 3x3 conv taps)?  Tries base_offset = 0 and
// base_offset = (addr >> 7) & 7 for shifts 0..9, 57, 58, 59, 117.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;

constexpr int ROWS = 256, K = 64, N = 64;

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int bo_mode) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    uint8_t* a_s = gen;                 // 256 rows x 128 B
    uint8_t* b_s = gen + ROWS * 128;    // 64 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < ROWS * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    for (int i = threadIdx.x; i < N * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<64>(smem_u32(&tslot));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a_addr = base + shift * 128;
        uint64_t adesc = make_smem_desc(a_addr, 0, 1024, 2);
        if (bo_mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
        const uint64_t bdesc = make_smem_desc(base + ROWS * 128, 0, 1024, 2);
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<64>(tm);
}

int main() {
    static __nv_bfloat16 hA[ROWS * K], hB[N * K];
    static float fA[ROWS * K], fB[N * K], hD[128 * N];
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const int smem = (ROWS + N) * 128 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int shifts[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 57, 58, 59, 117};
    for (int bo = 0; bo < 2; ++bo)
        for (int s : shifts) {
            cudaMemset(dD, 0, sizeof(hD));
            k<<<1, 128, smem>>>(dA, dB, dD, s, bo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("bo=%d shift=%d CUDA error %s\n", bo, s, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad_rows = 0;
            for (int j = 0; j < 128; ++j) {
                double rowerr = 0;
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int kk = 0; kk < K; ++kk) ref += (double)fA[(s + j) * K + kk] * fB[n * K + kk];
                    rowerr = fmax(rowerr, fabs(ref - hD[j * N + n]));
                }
                if (rowerr > 1e-3) ++bad_rows;
                maxerr = fmax(maxerr, rowerr);
            }
            printf("base_offset_mode=%d shift=%3d  max_err=%.5f bad_rows=%d\n", bo, s, maxerr, bad_rows);
        }
    return 0;
}
