// Experiment: tcgen05.mma with an MN-major B operand (V of attention: smem tile [key][d], d contiguous,
// 128B-swizzled as TMA writes it).  D[128 x 64] = P[128 x K] * V[K x 64], K = 128 keys.
// Tries (LBO, SBO) combinations for the B descriptor; A (= P) is K-major across two 64-wide atoms.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;
constexpr int M = 128, K = 128, N = 64;
__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* P, const __nv_bfloat16* V, float* D, int lbo, int sbo, int kstep) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    uint8_t* p_s = gen;               // two atoms: [128 rows x 64 keys] each, 16 KB apart
    uint8_t* v_s = gen + 32768;       // [128 keys][64 d] rows of 128 B, swizzled
    __shared__ uint64_t bar; __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < M * 16; i += 128) {           // P: row r, 16B chunk c (8 keys)
        int r = i >> 4, c = i & 15, atom = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(p_s + atom * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(P + r * K + c * 8);
    }
    for (int i = threadIdx.x; i < K * 8; i += 128) {            // V: key r, chunk c (8 d)
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(v_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(V + r * N + c * 8);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<64>(smem_u32(&tslot));
    fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(M, N, 0, 1);   // B is MN-major
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint32_t a_addr = base + (kk >> 2) * 16384 + (kk & 3) * 32;
            const uint64_t adesc = make_smem_desc(a_addr, 0, 1024, 2);
            const uint64_t bdesc = make_smem_desc(base + 32768 + kk * kstep, lbo, sbo, 2);
            umma_bf16(tm, adesc, bdesc, idesc, kk != 0);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0); tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<64>(tm);
}
int main() {
    static __nv_bfloat16 hP[M * K], hV[K * N]; static float fP[M * K], fV[K * N], hD[M * N];
    srand(2);
    for (int i = 0; i < M * K; ++i) { hP[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fP[i] = __bfloat162float(hP[i]); }
    for (int i = 0; i < K * N; ++i) { hV[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fV[i] = __bfloat162float(hV[i]); }
    __nv_bfloat16 *dP, *dV; float* dD;
    cudaMalloc(&dP, sizeof(hP)); cudaMalloc(&dV, sizeof(hV)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dP, hP, sizeof(hP), cudaMemcpyHostToDevice); cudaMemcpy(dV, hV, sizeof(hV), cudaMemcpyHostToDevice);
    const int smem = 32768 + 16384 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int cfgs[][3] = {{0, 1024, 2048}, {1024, 1024, 2048}, {1024, 0, 2048}, {128, 1024, 2048}, {1024, 128, 2048}, {2048, 1024, 2048}, {1024, 2048, 2048}, {0, 2048, 2048}};
    for (auto& c : cfgs) {
        cudaMemset(dD, 0, sizeof(hD));
        k<<<1, 128, smem>>>(dP, dV, dD, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo=%d sbo=%d: CUDA error %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int kk = 0; kk < K; ++kk) ref += (double)fP[m * K + kk] * fV[kk * N + n];
            maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
        }
        printf("lbo=%4d sbo=%4d kstep=%d  max_err=%.5f %s\n", c[0], c[1], c[2], maxerr, maxerr < 1e-2 ? "OK" : "");
    }
    return 0;
}
