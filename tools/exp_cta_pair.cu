// Experiment: tcgen05.mma.cta_group::2 - one 256 x 256 x 64 product issued by the leader CTA of a two-CTA cluster.
// Each CTA holds 128 rows of A and 128 of the 256 rows of B (K-major, SWIZZLE_128B) at the same shared-memory
// offsets; each CTA's TMEM receives its own 128 accumulator rows x 256 columns.  Checks: who allocates TMEM, the
// instruction descriptor (M = 256), the multicast commit, and that the peer's operands are read from the peer's smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/exp_cta_pair tools/exp_cta_pair.cu && /tmp/exp_cta_pair
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;

constexpr int M = 256, N = 256, K = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    uint8_t* a_s = gen;                 // 128 rows x 128 B: A rows [128 r, 128 r + 128)
    uint8_t* b_s = gen + 128 * 128;     // 128 rows x 128 B: B rows [128 r, 128 r + 128)
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(A + (rank * 128 + r) * K + c * 8);
        *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(B + (rank * 128 + r) * K + c * 8);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) {   // one warp of EACH CTA of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (rank == 0 && threadIdx.x == 0) {
        const uint64_t adesc = make_smem_desc(base, 0, 1024, 2);
        const uint64_t bdesc = make_smem_desc(base + 128 * 128, 0, 1024, 2);
        const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
        for (int kk = 0; kk < 4; ++kk) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                "l"(adesc + 2 * kk), "l"(bdesc + 2 * kk), "r"(idesc), "r"(kk != 0 ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                         "r"(smem_u32(&bar)), "h"(static_cast<uint16_t>(3))
                     : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(rank * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256) : "memory");
}

int main() {
    static __nv_bfloat16 hA[M * K], hB[N * K];
    static float fA[M * K], fB[N * K], hD[M * N];
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const int smem = 2 * 128 * 128 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(dD, 0xff, sizeof(hD));
        k<<<2, 128, smem>>>(dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < K; ++kk) ref += (double)fA[m * K + kk] * fB[n * K + kk];
                const double err = fabs(ref - hD[m * N + n]);
                if (!(err <= 1e-3)) { if (bad < 5) printf("  bad (%d,%d): got %f want %f\n", m, n, hD[m * N + n], ref); ++bad; }
                if (err == err) maxerr = fmax(maxerr, err);
            }
        printf("rep %d: max_err=%.6f bad=%d of %d\n", rep, maxerr, bad, M * N);
    }
    return 0;
}
