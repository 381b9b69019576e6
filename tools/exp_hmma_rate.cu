// Microbenchmark: peak throughput of the legacy mma.sync.m16n8k16 bf16 path on sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float d[8][4];
    for (int j = 0; j < 8; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
    uint32_t a[4] = {threadIdx.x, 1u, 2u, 3u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    for (int bps : {1, 2, 4, 8}) {
        const int iters = 20000;
        k<<<148 * bps, 256>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148 * bps, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = 148.0 * bps * 8 * iters * 8;       // warps * iters * 8
        double tflops = mmas * 16 * 8 * 16 * 2 / (ms * 1e-3) / 1e12;
        printf("blocks/SM=%d warps/SM=%d: %.3f ms, %.1f TFLOP/s (dense bf16 mma.sync m16n8k16)\n", bps, bps * 8, ms, tflops);
    }
    return 0;
}
