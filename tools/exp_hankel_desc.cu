// Experiment: can an un-swizzled K-major tcgen05 A descriptor describe OVERLAPPING rows - row m starting 16 bytes
// after row m-1 inside an 8-row group (the canonical core matrix), K chunk c another 16 bytes further (LBO = 16 B,
// i.e. core matrices adjacent in K overlap by 7/8), 8-row groups SBO bytes apart?  That is the im2col matrix of a
// stride-2 7x7 stem read straight out of the raw input rows (8 output pixels of one row = one core-matrix group,
// output rows = groups 2 input rows apart), with no im2col copy at all.
// Tries both assignments of (LBO, SBO) to (K direction, M direction).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;

constexpr int N = 64, KT = 32;            // two K = 16 steps
constexpr int RAW = 16 * 384 + 512;       // 16 groups x 384 B + slack

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* raw_g, const __nv_bfloat16* B, float* D, int lbo, int sbo) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* b_s = gen + 8192;            // 64 rows x 128 B, SWIZZLE_128B
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < RAW / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(gen)[i] = raw_g[i];
    for (int i = threadIdx.x; i < N * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<64>(smem_u32(&tslot));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        for (int kk = 0; kk < KT / 16; ++kk) {
            const uint64_t adesc = make_smem_desc(base + kk * 32, lbo, sbo, 0);   // no swizzle
            const uint64_t bdesc = make_smem_desc(base + 8192, 0, 1024, 2) + 2 * kk;
            umma_bf16(tm, adesc, bdesc, idesc, kk != 0);
        }
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
    static __nv_bfloat16 hR[RAW / 2], hB[N * 64];
    static float fR[RAW / 2], fB[N * 64], hD[128 * N];
    srand(2);
    for (int i = 0; i < RAW / 2; ++i) { hR[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fR[i] = __bfloat162float(hR[i]); }
    for (int i = 0; i < N * 64; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dR, *dB; float* dD;
    cudaMalloc(&dR, sizeof(hR)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dR, hR, sizeof(hR), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const int smem = 8192 + N * 128 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // hypotheses: (stride between K chunks, stride between 8-row groups)
    const int kstr = 16, mstr = 384;
    const int trials[4][2] = {{kstr, mstr}, {mstr, kstr}, {16, 128}, {128, 16}};
    for (auto& t : trials) {
        cudaMemset(dD, 0, sizeof(hD));
        k<<<1, 128, smem>>>(dR, dB, dD, t[0], t[1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo=%d sbo=%d CUDA error %s\n", t[0], t[1], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        // model A: K-chunk stride = LBO, group stride = SBO ; model B: the other way round
        for (int model = 0; model < 2; ++model) {
            const int ks = model == 0 ? t[0] : t[1], ms = model == 0 ? t[1] : t[0];
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int kk = 0; kk < KT; ++kk) {
                        const int byte = (kk / 16) * 32 + (m / 8) * ms + (m % 8) * 16 + ((kk % 16) / 8) * ks + (kk % 8) * 2;
                        ref += (double)fR[byte / 2] * fB[n * 64 + kk];
                    }
                    maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
                }
            printf("lbo=%3d sbo=%3d  model %s: max_err=%.5f %s\n", t[0], t[1],
                   model == 0 ? "K-stride=LBO, group-stride=SBO" : "K-stride=SBO, group-stride=LBO", maxerr,
                   maxerr < 1e-3 ? "MATCH" : "");
        }
    }
    return 0;
}
