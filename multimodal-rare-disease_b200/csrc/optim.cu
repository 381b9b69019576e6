// Optimizer step of the training configuration (reference: src/train.py:307-320 -
// clip_grad_norm_(parameters, 1.0) then torch.optim.AdamW.step()): ONE multi-tensor pass for the global gradient
// norm and ONE for clip + AdamW over every parameter, no host synchronisation in between (the clip coefficient is
// computed on the device from the reduced norm).  fp32 parameters, gradients and moments; HBM-bound:
// 4 reads + 3 writes of 4 bytes per element = 28 B/element (113 M trainable elements -> 3.2 GB per step).

#include <mrd_b200.h>

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tma_host.h"

namespace {

struct AdamTensor {   // mirrors mrd_adamw_tensor (include/mrd_b200.h)
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    float lr;
    float wd;
};
static_assert(sizeof(AdamTensor) == sizeof(mrd_adamw_tensor), "layout");

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x < (blockDim.x >> 5)) t = red[threadIdx.x];
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in thread 0
}

// chunk c covers elements [chunk_off[c], chunk_off[c] + chunk_elems) of tensor chunk_tensor[c]
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const AdamTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
                   const long long* __restrict__ chunk_off, int chunk_elems, float* __restrict__ sqnorm) {
    __shared__ float red[8];
    const AdamTensor t = tensors[chunk_tensor[blockIdx.x]];
    const long long lo = chunk_off[blockIdx.x];
    const long long hi = lo + chunk_elems < t.n ? lo + chunk_elems : t.n;
    float acc = 0.0f;
    if (t.g) {
        if ((reinterpret_cast<uintptr_t>(t.g) & 15) == 0) {
            const long long lo4 = lo / 4, hi4 = hi / 4;   // chunk boundaries are multiples of 4 except the tensor's tail
            for (long long i = lo4 + threadIdx.x; i < hi4; i += blockDim.x) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(t.g) + i);
                acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
            }
            for (long long i = hi4 * 4 + threadIdx.x; i < hi; i += blockDim.x) acc += t.g[i] * t.g[i];
        } else {
            for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += t.g[i] * t.g[i];
        }
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0 && s != 0.0f) atomicAdd(sqnorm, s);
}

__global__ void __launch_bounds__(256)
adamw_kernel(const AdamTensor* __restrict__ tensors, const int* __restrict__ chunk_tensor,
             const long long* __restrict__ chunk_off, int chunk_elems, float beta1, float beta2, float eps,
             float bias1, float bias2_sqrt, float max_norm, const float* __restrict__ sqnorm) {
    const AdamTensor t = tensors[chunk_tensor[blockIdx.x]];
    if (!t.g) return;   // no gradient this step: torch skips the parameter entirely
    const long long lo = chunk_off[blockIdx.x];
    const long long hi = lo + chunk_elems < t.n ? lo + chunk_elems : t.n;
    float coef = 1.0f;
    if (max_norm > 0.0f) {   // clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max = 1)
        coef = max_norm / (sqrtf(__ldg(sqnorm)) + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;
    }
    const float decay = 1.0f - t.lr * t.wd;
    const float step_size = t.lr / bias1;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float g = t.g[i] * coef;
        float p = t.p[i] * decay;
        const float m = beta1 * t.m[i] + (1.0f - beta1) * g;
        const float v = beta2 * t.v[i] + (1.0f - beta2) * g * g;
        t.m[i] = m;
        t.v[i] = v;
        p -= step_size * m / (sqrtf(v) / bias2_sqrt + eps);
        t.p[i] = p;
    }
}

}  // namespace

extern "C" int mrd_adamw_step(const mrd_adamw_tensor* tensors_dev, int n_tensors, const int* chunk_tensor_dev,
                              const long long* chunk_off_dev, int n_chunks, int chunk_elems, float beta1,
                              float beta2, float eps, long long step, float max_norm, float* sqnorm_dev,
                              void* stream) {
    if (n_tensors <= 0 || n_chunks <= 0) return 0;
    if (!tensors_dev || !chunk_tensor_dev || !chunk_off_dev || !sqnorm_dev || chunk_elems <= 0 || chunk_elems % 4 ||
        step <= 0) {
        mrd::set_last_error("mrd_adamw_step: bad arguments");
        return -1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const AdamTensor* t = reinterpret_cast<const AdamTensor*>(tensors_dev);
    cudaError_t e = cudaMemsetAsync(sqnorm_dev, 0, sizeof(float), s);
    if (e != cudaSuccess) {
        mrd::set_last_error("mrd_adamw_step: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    grad_sqnorm_kernel<<<n_chunks, 256, 0, s>>>(t, chunk_tensor_dev, chunk_off_dev, chunk_elems, sqnorm_dev);
    const float bias1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bias2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    adamw_kernel<<<n_chunks, 256, 0, s>>>(t, chunk_tensor_dev, chunk_off_dev, chunk_elems, beta1, beta2, eps, bias1,
                                          bias2_sqrt, max_norm, sqnorm_dev);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        mrd::set_last_error("mrd_adamw_step launch: %s", cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}
