// Microbenchmark: completion rate of tcgen05.mma.cta_group::2 (256 x N x 16 over a CTA pair) against
// cta_group::1 (128 x N x 16 per CTA), operands in shared memory (K-major, SWIZZLE_128B), 4 operand buffers cycled.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;

template <int N, int CG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(long long* out, int iters, int commit_each) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[8];   // per-step commits land here (counts are large enough never to complete a phase)
    __shared__ uint32_t tslot;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1 << 20);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) {
        if (CG == 2) tmem_alloc_pair<256>(smem_u32(&tslot)); else tmem_alloc<256>(smem_u32(&tslot));
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0 && (CG == 1 || rank == 0)) {
        const uint32_t idesc = make_idesc_bf16(CG == 2 ? 256 : 128, N, 0, 0);
        const int brows = CG == 2 ? N / 2 : N;      // rows of B in this CTA's shared memory
        const uint32_t b_addr = base + 64 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t adesc = make_smem_desc(base + (i % 4) * 16384, 0, 1024, 2);
            const uint64_t bdesc = make_smem_desc(b_addr + (i % 4) * brows * 128, 0, 1024, 2);
            for (int kk = 0; kk < 4; ++kk) {
                if (CG == 2) umma_bf16_pair(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1);
                else umma_bf16(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1);
            }
            if (commit_each) {
                if (CG == 2) umma_commit_pair(smem_u32(&dummy[i % 8]), 3); else umma_commit(smem_u32(&dummy[i % 8]));
            }
        }
        long long t1 = clock64();
        if (CG == 2) umma_commit_pair(smem_u32(&bar), 1); else umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[rank * 2 + 0] = t1 - t0; out[rank * 2 + 1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) {
        if (CG == 2) tmem_dealloc_pair<256>(tm); else tmem_dealloc<256>(tm);
    }
}
template <int N, int CG> void run(long long* d, int commit_each = 0) {
    const int iters = 900, smem = 200 * 1024;
    cudaFuncSetAttribute(k<N, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, CG><<<2, 128, smem>>>(d, iters, commit_each); cudaDeviceSynchronize();
    k<N, CG><<<2, 128, smem>>>(d, iters, commit_each);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("cta_group::%d N=%3d commit_each=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%.0f MACs/clk/SM) %s\n", CG, N, commit_each,
           h[0] / (4.0 * iters), h[1] / (4.0 * iters), 128.0 * N * 16 / (h[1] / (4.0 * iters)),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    long long* d; cudaMalloc(&d, 32);
    run<256, 1>(d); run<256, 2>(d); run<128, 1>(d); run<128, 2>(d); run<64, 2>(d);
    run<256, 1>(d, 1); run<256, 2>(d, 1);
    return 0;
}
