// Strided SIMT GEMM (fp32 accumulate, plain FFMA): C[M,N] = act(alpha * A * B^T + bias) (+ residual).
//
// Not the hot path - the tcgen05 kernel in gemm_conv.cu is.  This kernel exists for
//   * the fp32 check mode (fp32_check.cu): every contraction of the forward in full fp32, including the
//     convolutions (implicit GEMM straight from NCHW, un-folded BatchNorm in the epilogue), so the
//     whole pipeline can be compared with the reference at 1e-4 (BASELINE.json north_star);
//   * the small batch-level layers of the training step (train.cu): M = batch size rows, where a
//     128-row tensor-core tile would be mostly padding, and where operands are needed transposed
//     (dX = dY W, dW = dY^T X) - arbitrary element strides make every variant the same kernel.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mrd {

struct SimtConv {        // A is an NCHW fp32 image tensor; M = Nimg*Ho*Wo, K = Cin*ks*ks
    int Cin, H, W, Ho, Wo, ks, stride, pad;
    // eval-mode BatchNorm applied to the accumulator exactly as written in the reference stack
    // (TV:models/resnet.py:143-163): (x - mean) / sqrt(var + eps) * gamma + beta
    const float *mean, *var, *gamma, *beta;
    float eps;
    const float* residual;   // NCHW fp32, same shape as the output, or null
};

struct SimtGemm {
    const void* A = nullptr;   // A(m,k) = A[m*a_rs + k*a_cs]
    long long a_rs = 0, a_cs = 1;
    int a_bf16 = 0;
    const void* B = nullptr;   // B(n,k) = B[n*b_rs + k*b_cs]  (nn.Linear weight [N,K]: b_rs = K, b_cs = 1)
    long long b_rs = 0, b_cs = 1;
    int b_bf16 = 0;
    float* C = nullptr;        // fp32 output, row stride ldc (optional)
    long long ldc = 0;
    __nv_bfloat16* C16 = nullptr;  // bf16 output, row stride ldc16 (optional)
    long long ldc16 = 0;
    const float* bias = nullptr;   // [N]
    const float* res = nullptr;    // fp32 residual [M,N], row stride ldr, added after the activation
    long long ldr = 0;
    float alpha = 1.0f;
    int act = 0;               // MRD_ACT_*
    int accumulate = 0;        // C += result (fp32 output only)
    int M = 0, N = 0, K = 0;
    int ksplit = 0;            // 0: automatic.  > 1: the K loop is cut into slices (grid.z) whose partial sums are
                               // added atomically into C (fp32 only, no activation; bias / residual come from
                               // slice 0); C is zeroed first unless `accumulate` is set.  Skinny problems
                               // (rows = batch size) otherwise run a long serial K loop on a handful of CTAs.
    const int* dyn_k = nullptr;    // optional device int: contraction length min(K, *dyn_k)
    const int* dyn_m = nullptr;    // optional device int: rows min(M, *dyn_m)
};

int simt_gemm(const SimtGemm& g, cudaStream_t s);
// Convolution + BatchNorm(eval) + residual + activation, NCHW fp32 in / out.  Wt: [Cout, Cin*ks*ks].
int simt_conv_bn(const float* x, int Nimg, const SimtConv& cv, const float* Wt, int Cout, int act,
                 float* y, cudaStream_t s);

}  // namespace mrd
