// Microbenchmark: tcgen05.mma issue/throughput for M=128, N in {64,128,256}, K=16 with the A descriptor
// starting at a 1024-aligned address vs at a row-shifted (128 B granular) address.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int shift, int iters, int same_desc) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tslot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t b_addr = base + 64 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int tap = same_desc ? 0 : (i % 9);
            const uint32_t a_addr = base + (shift + (tap / 3) * 58 + tap % 3) * 128;
            const uint64_t adesc = make_smem_desc(a_addr, 0, 1024, 2);
            const uint64_t bdesc = make_smem_desc(b_addr + (i % 4) * N * 128, 0, 1024, 2);
            for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tm);
}
template <int N> void run(long long* d, int shift, int same) {
    const int iters = 900, smem = 200 * 1024;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N><<<1, 128, smem>>>(d, shift, iters, same); cudaDeviceSynchronize();
    k<N><<<1, 128, smem>>>(d, shift, iters, same);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d shift=%2d same_desc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d) %s\n", N, shift, same,
           h[0] / (4.0 * iters), h[1] / (4.0 * iters), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    for (int same : {1, 0}) for (int shift : {0, 1, 8}) { run<64>(d, shift, same); }
    run<128>(d, 0, 1); run<128>(d, 1, 0); run<256>(d, 0, 1);
    return 0;
}
