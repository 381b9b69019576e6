// Microbenchmark: does streaming operands into shared memory (cp.async.bulk, the TMA engine) slow down tcgen05.mma
// that reads its operands from the same shared memory?  One CTA per SM on every SM; thread 0 issues 128 x 256 x 16
// MMAs back to back, thread 32 keeps `depth` bulk copies of `chunk` bytes in flight from a large global buffer.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../multimodal-rare-disease_b200/csrc/ptx.cuh"
using namespace mrd;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t src_bytes, long long* out, int iters, int copy_on,
                                            int mma_on, int chunk) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar, cbar[4];
    __shared__ uint32_t tslot;
    __shared__ volatile int done;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1);
        fence_mbar_init();
        done = 0;
    }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tslot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tslot;
    const uint32_t copy_base = base + 96 * 1024;   // 4 x 32 KB ring above the 96 KB the MMAs read
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
        long long t0 = clock64();
        if (mma_on) {
            for (int i = 0; i < iters; ++i) {
                const uint64_t adesc = make_smem_desc(base + (i % 2) * 16384, 0, 1024, 2);
                const uint64_t bdesc = make_smem_desc(base + 32768 + (i % 2) * 32768, 0, 1024, 2);
                for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1);
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
        } else {
            while (clock64() - t0 < 400000) {}
        }
        long long t1 = clock64();
        done = 1;
        out[blockIdx.x * 4 + 0] = t1 - t0;
    } else if (threadIdx.x == 32 && copy_on) {
        long long t0 = clock64();
        long long bytes = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        size_t off = (size_t)blockIdx.x * 1048576 % src_bytes;
        for (int i = 0; i < 4; ++i) {
            mbar_expect_tx(smem_u32(&cbar[i]), chunk);
            bulk_g2s(copy_base + i * 32768, src + off, chunk, smem_u32(&cbar[i]));
            off = (off + chunk) % (src_bytes - 65536);
        }
        int i = 0;
        while (!done) {
            mbar_wait(smem_u32(&cbar[i]), ph[i]);
            ph[i] ^= 1;
            bytes += chunk;
            mbar_expect_tx(smem_u32(&cbar[i]), chunk);
            bulk_g2s(copy_base + i * 32768, src + off, chunk, smem_u32(&cbar[i]));
            off = (off + chunk) % (src_bytes - 65536);
            i = (i + 1) & 3;
        }
        for (int j = 0; j < 4; ++j) { mbar_wait(smem_u32(&cbar[i]), ph[i]); i = (i + 1) & 3; }
        long long t1 = clock64();
        out[blockIdx.x * 4 + 1] = bytes;
        out[blockIdx.x * 4 + 2] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tm);
}

int main() {
    const size_t src_bytes = 64ull << 20;   // L2-resident source
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    long long* d; cudaMalloc(&d, 148 * 4 * 8);
    const int smem = 225 * 1024 + 512;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 800;
    for (int chunk : {16384, 32768})
        for (int mode = 0; mode < 3; ++mode) {
            const int copy_on = mode != 0, mma_on = mode != 2;
            cudaMemset(d, 0, 148 * 4 * 8);
            k<<<148, 128, smem>>>(src, src_bytes, d, iters, copy_on, mma_on, chunk);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            long long h[148 * 4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double mma = 0, bpc = 0;
            for (int b = 0; b < 148; ++b) { mma += h[b * 4] / (4.0 * iters); if (h[b * 4 + 2]) bpc += (double)h[b * 4 + 1] / h[b * 4 + 2]; }
            printf("chunk %5d  mma %d copy %d: %.1f cyc/MMA (floor 128), copy stream %.1f B/clk/SM  %s\n", chunk, mma_on, copy_on,
                   mma_on ? mma / 148 : 0.0, bpc / 148, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
