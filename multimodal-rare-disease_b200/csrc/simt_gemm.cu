// Strided SIMT GEMM / implicit-GEMM convolution in fp32 (see simt_gemm.h).

#include "simt_gemm.h"

#include <math.h>

#include <mrd_b200.h>

#include "tma_host.h"

namespace mrd {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ float ld_elem(const void* p, long long i, int is_bf16) {
    return is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i])
                   : __ldg(static_cast<const float*>(p) + i);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == MRD_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == MRD_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    return v;
}

template <bool CONV>
__global__ void __launch_bounds__(256) simt_gemm_kernel(SimtGemm p, SimtConv cv) {
    __shared__ float As[BK][BM + PAD];
    __shared__ float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    // conv output is NCHW (m contiguous): let the fast thread index walk m there, n otherwise
    const int tm = CONV ? (tid & 15) : (tid >> 4);
    const int tn = CONV ? (tid >> 4) : (tid & 15);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    int K = p.K, M = p.M;
    if (p.dyn_k) K = min(K, *p.dyn_k);
    if (p.dyn_m) M = min(M, *p.dyn_m);
    if (m0 >= M) return;
    const bool a_kfast = !CONV && p.a_cs == 1;
    const bool b_kfast = p.b_cs == 1;
    const int HoWo = CONV ? cv.Ho * cv.Wo : 1;
    const int kk2 = CONV ? cv.ks * cv.ks : 1;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    int k_lo = 0;
    if (!CONV && gridDim.z > 1) {   // split-K: this CTA's slice of the contraction
        const int per = ((K + static_cast<int>(gridDim.z) - 1) / static_cast<int>(gridDim.z) + BK - 1) / BK * BK;
        k_lo = static_cast<int>(blockIdx.z) * per;
        K = min(K, k_lo + per);
    }
    for (int k0 = k_lo; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int mm, kk;
            if (a_kfast) { kk = e & (BK - 1); mm = e >> 4; } else { mm = e & (BM - 1); kk = e >> 6; }
            const int m = m0 + mm, k = k0 + kk;
            float v = 0.0f;
            if (m < M && k < K) {
                if (CONV) {
                    const int img = m / HoWo, hw = m - img * HoWo;
                    const int ho = hw / cv.Wo, wo = hw - ho * cv.Wo;
                    const int ci = k / kk2, r = k - ci * kk2;
                    const int kh = r / cv.ks, kw = r - kh * cv.ks;
                    const int hi = ho * cv.stride - cv.pad + kh, wi = wo * cv.stride - cv.pad + kw;
                    if (hi >= 0 && hi < cv.H && wi >= 0 && wi < cv.W)
                        v = __ldg(static_cast<const float*>(p.A) +
                                  ((static_cast<long long>(img) * cv.Cin + ci) * cv.H + hi) * cv.W + wi);
                } else {
                    v = ld_elem(p.A, m * p.a_rs + k * p.a_cs, p.a_bf16);
                }
            }
            As[kk][mm] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int nn, kk;
            if (b_kfast) { kk = e & (BK - 1); nn = e >> 4; } else { nn = e & (BN - 1); kk = e >> 6; }
            const int n = n0 + nn, k = k0 + kk;
            float v = 0.0f;
            if (n < p.N && k < K) v = ld_elem(p.B, n * p.b_rs + k * p.b_cs, p.b_bf16);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tm * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tn * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j] * p.alpha;
            if (p.bias && gridDim.z == 1) v += __ldg(p.bias + n);
            if (CONV) {
                v = (v - __ldg(cv.mean + n)) / sqrtf(__ldg(cv.var + n) + cv.eps) * __ldg(cv.gamma + n) +
                    __ldg(cv.beta + n);
                const int img = m / HoWo, hw = m - img * HoWo;
                const long long o = (static_cast<long long>(img) * p.N + n) * HoWo + hw;
                if (cv.residual) v += __ldg(cv.residual + o);
                p.C[o] = apply_act(v, p.act);
            } else if (gridDim.z > 1) {
                float v2 = acc[i][j] * p.alpha;
                if (blockIdx.z == 0) {
                    if (p.bias) v2 += __ldg(p.bias + n);
                    if (p.res) v2 += __ldg(p.res + m * p.ldr + n);
                }
                atomicAdd(p.C + m * p.ldc + n, v2);
            } else {
                v = apply_act(v, p.act);
                if (p.res) v += __ldg(p.res + m * p.ldr + n);
                if (p.C) {
                    float* c = p.C + m * p.ldc + n;
                    *c = p.accumulate ? *c + v : v;
                }
                if (p.C16) p.C16[m * p.ldc16 + n] = __float2bfloat16(v);
            }
        }
    }
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s launch: %s", what, cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

}  // namespace

int simt_gemm(const SimtGemm& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (!g.A || !g.B || (!g.C && !g.C16)) {
        set_last_error("simt_gemm: null operand");
        return -1;
    }
    dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
    int ks = g.ksplit;
    const bool can_split = g.C && !g.C16 && g.act == MRD_ACT_NONE && !g.dyn_k && !g.dyn_m &&
                           (g.accumulate || g.ldc == g.N);
    if (ks == 0 && can_split) {
        const long long tiles = static_cast<long long>(grid.x) * grid.y;
        if (tiles < 96 && g.K >= 256) {
            ks = static_cast<int>(148 / tiles);
            if (ks > g.K / 128) ks = g.K / 128;
        }
    }
    if (ks > 1 && can_split) {
        if (!g.accumulate) {
            cudaError_t e = cudaMemsetAsync(g.C, 0, sizeof(float) * static_cast<size_t>(g.M) * g.N, s);
            if (e != cudaSuccess) {
                set_last_error("simt_gemm: cudaMemsetAsync: %s", cudaGetErrorString(e));
                return -static_cast<int>(e);
            }
        }
        grid.z = ks;
    }
    SimtConv none{};
    simt_gemm_kernel<false><<<grid, 256, 0, s>>>(g, none);
    return check_launch("simt_gemm");
}

int simt_conv_bn(const float* x, int Nimg, const SimtConv& cv, const float* Wt, int Cout, int act,
                 float* y, cudaStream_t s) {
    SimtGemm g;
    g.A = x;
    g.B = Wt;
    g.K = cv.Cin * cv.ks * cv.ks;
    g.b_rs = g.K;
    g.b_cs = 1;
    g.M = Nimg * cv.Ho * cv.Wo;
    g.N = Cout;
    g.C = y;
    g.act = act;
    dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
    simt_gemm_kernel<true><<<grid, 256, 0, s>>>(g, cv);
    return check_launch("simt_conv_bn");
}

}  // namespace mrd
