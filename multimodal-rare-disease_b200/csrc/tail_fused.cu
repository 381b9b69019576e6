// tail_fused_kernel (K6): AttentionFusion + ClassificationHead + softmax for 64 samples per CTA, one launch
// (see tail_fused.h for the algebra).
//
//   warp 0          TMA producer   walks the stage list: per accumulator tile and 64-wide K chunk one ring slot =
//                                  [64 x 64 embedding chunk (first two stages only)][bn x 64 weight tile]
//   warp 1          MMA issuer     128 x bn x 16 tcgen05.mma per K slice into TMEM columns [0, n); owns TMEM.
//                                  The A operand has 64 real rows: its 8 KB chunks sit back to back, so the upper 64
//                                  rows the instruction reads are the next chunk's bytes - they only produce
//                                  accumulator lanes 64..127, which nobody reads (an M = 64 instruction occupies the
//                                  tensor core just as long as an M = 128 one)
//   warps 4,5,8,9   epilogue       TMEM lane quarters 0 and 1 (= the 64 real rows), two warps per quarter split the
//                                  columns: +bias, then LayerNorm (two-pass statistics over the whole row, halves
//                                  exchanged through shared memory) or ReLU/GELU -> bf16 -> next stage's A operand;
//                                  last stage: GEMV with the final Linear + softmax -> logits / probabilities
// Stages are strictly sequential (each consumes the previous one's full output), so one accumulator and one
// "operand ready" barrier suffice; the producer prefetches the next stage's weights while an epilogue runs.
// Reference ops replaced: src/fusion_model.py:116-182,245-291, src/multimodal_classifier.py:73-83,166-167.

#include "tail_fused.h"

#include <stdio.h>
#include <string.h>

#include "gemm_conv.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace mrd {

namespace {

constexpr int kRows = 64;                  // samples per CTA
constexpr int kChunk = kRows * 128;        // 64 rows x 64 bf16 columns, SWIZZLE_128B K-major: 8 KB
constexpr int kActChunks = 16;             // resident activation area: up to 1024 columns
constexpr int kWTile = 256 * 128;          // weight tile: up to 256 rows x 64 K
constexpr int kSlot = kChunk + kWTile;     // ring slot: streamed embedding chunk + weight tile
constexpr int kPipe = 2;
constexpr int kThreads = 320;
constexpr int kEpi = 128;
constexpr int kMaxC = 32;
constexpr int kMisc = 1024 + kRows * kMaxC * 4;   // LayerNorm statistics exchange + GEMV partial sums
constexpr int kSmem = 1024 + kActChunks * kChunk + kPipe * kSlot + kMisc + 256;

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_RELU) return fmaxf(x, 0.0f);
    if (act == ACT_GELU) return gelu_erf_fast(x);
    return x;
}

__global__ void __launch_bounds__(kThreads, 1) tail_fused_kernel(const __grid_constant__ TailParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t act_s = base;
    const uint32_t ring_s = act_s + kActChunks * kChunk;
    const uint32_t misc_s = ring_s + kPipe * kSlot;
    const uint32_t bar = misc_s + kMisc;
    uint8_t* act_gen = gen;
    float* stats = reinterpret_cast<float*>(gen + (misc_s - base));   // [2][64][2]: sums, then squared deviations
    float* part = stats + 256;                                         // [64][kMaxC]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar - base) + 96);

    auto full = [&](int s) { return bar + 8u * s; };
    auto empty = [&](int s) { return bar + 32u + 8u * s; };
    const uint32_t tfull = bar + 64u, aready = bar + 72u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.img_map);
        tma_prefetch_desc(&p.txt_map);
        for (int s = 0; s < p.num_stages; ++s) tma_prefetch_desc(&p.st[s].w_map);
        for (int s = 0; s < kPipe; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tfull, 1);
        mbar_init(aready, kEpi);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(bar + 96);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int row0 = static_cast<int>(blockIdx.x) * kRows;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        int slot = 0;
        uint32_t ph = 0;
        for (int s = 0; s < p.num_stages; ++s) {
            const TailStage& S = p.st[s];
            const int ntiles = S.n / S.bn;
            const bool streamed = S.a_chunk0 < 0;
            for (int nt = 0; nt < ntiles; ++nt)
                for (int kc = 0; kc < S.k_chunks; ++kc) {
                    mbar_wait(empty(slot), ph ^ 1u);
                    if (lane == 0) {
                        const uint32_t dst = ring_s + slot * kSlot;
                        mbar_expect_tx(full(slot), static_cast<uint32_t>(S.bn * 128 + (streamed ? kChunk : 0)));
                        if (streamed) {
                            if (kc < p.img_chunks)
                                tma_load_2d(&p.img_map, full(slot), dst, kc * 64, row0);
                            else
                                tma_load_2d(&p.txt_map, full(slot), dst, (kc - p.img_chunks) * 64, row0);
                        }
                        tma_load_2d(&S.w_map, full(slot), dst + kChunk, kc * 64, S.w_row0 + nt * S.bn);
                    }
                    __syncwarp();
                    if (++slot == kPipe) { slot = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        int slot = 0;
        uint32_t ph = 0;
        for (int s = 0; s < p.num_stages; ++s) {
            const TailStage& S = p.st[s];
            const int ntiles = S.n / S.bn;
            const bool streamed = S.a_chunk0 < 0;
            // the previous stage's epilogue has drained the accumulator and written this stage's A operand
            if (s > 0) mbar_wait(aready, static_cast<uint32_t>(s - 1) & 1u);
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, S.bn, 0, 0);
            for (int nt = 0; nt < ntiles; ++nt)
                for (int kc = 0; kc < S.k_chunks; ++kc) {
                    mbar_wait(full(slot), ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sbase = ring_s + slot * kSlot;
                        const uint32_t a_addr = streamed ? sbase : act_s + (S.a_chunk0 + kc) * kChunk;
                        const uint64_t ad = make_smem_desc(a_addr, 0, 1024, 2);
                        const uint64_t bd = make_smem_desc(sbase + kChunk, 0, 1024, 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem + nt * S.bn, ad + 2 * k, bd + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                        umma_commit(empty(slot));
                        if (nt == ntiles - 1 && kc == S.k_chunks - 1) umma_commit(tfull);
                    }
                    __syncwarp();
                    if (++slot == kPipe) { slot = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 4 || warp == 5 || warp == 8 || warp == 9) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;            // TMEM lane quarter 0 / 1
        const int h = warp >> 3;           // column half
        const int row = q * 32 + lane;     // 0..63
        const int grow = row0 + row;
        const bool valid = grow < p.B;
        uint32_t v[32];
        for (int s = 0; s < p.num_stages; ++s) {
            const TailStage& S = p.st[s];
            const int half = S.n >> 1;
            const int c0 = h * half;
            const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + c0;
            const float inv_n = 1.0f / static_cast<float>(S.n);
            mbar_wait(tfull, static_cast<uint32_t>(s) & 1u);
            tc_fence_after();

            // bf16 row segment [col, col + 32) of the stage result -> A operand of a later stage
            auto store_bf16 = [&](int col, const float* y) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int cc = col + u * 8;
                    uint4 o;
                    o.x = pack_bf16(y[u * 8 + 0], y[u * 8 + 1]);
                    o.y = pack_bf16(y[u * 8 + 2], y[u * 8 + 3]);
                    o.z = pack_bf16(y[u * 8 + 4], y[u * 8 + 5]);
                    o.w = pack_bf16(y[u * 8 + 6], y[u * 8 + 7]);
                    uint8_t* dst = act_gen + (S.out_chunk0 + (cc >> 6)) * kChunk + row * 128 +
                                   ((((cc & 63) >> 3) ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = o;
                }
            };
            auto load_biased = [&](int j, float* x) {   // accumulator columns [c0 + 32 j, +32) plus bias
                tmem_ld32(taddr + j * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(S.bias + c0 + j * 32 + i));
                    x[i + 0] = __uint_as_float(v[i + 0]) + b4.x;
                    x[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
                    x[i + 2] = __uint_as_float(v[i + 2]) + b4.z;
                    x[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
                }
            };
            float x[32];
            if (S.epi == TAIL_EPI_LN) {
                // nn.LayerNorm over the whole row (src/fusion_model.py:270,276): mean, then squared deviations
                float sum = 0.0f;
                for (int j = 0; j < half / 32; ++j) {
                    load_biased(j, x);
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum += x[i];
                }
                stats[row * 2 + h] = sum;
                named_bar_sync(1, kEpi);
                const float mean = (stats[row * 2] + stats[row * 2 + 1]) * inv_n;
                float ssq = 0.0f;
                for (int j = 0; j < half / 32; ++j) {
                    load_biased(j, x);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float d = x[i] - mean;
                        ssq = fmaf(d, d, ssq);
                    }
                }
                stats[128 + row * 2 + h] = ssq;
                named_bar_sync(1, kEpi);
                const float rstd = rsqrtf((stats[128 + row * 2] + stats[128 + row * 2 + 1]) * inv_n + p.ln_eps);
                for (int j = 0; j < half / 32; ++j) {
                    load_biased(j, x);
                    const int col = c0 + j * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(S.ln_g + col + i));
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(S.ln_b + col + i));
                        x[i + 0] = fmaf((x[i + 0] - mean) * rstd, g4.x, b4.x);
                        x[i + 1] = fmaf((x[i + 1] - mean) * rstd, g4.y, b4.y);
                        x[i + 2] = fmaf((x[i + 2] - mean) * rstd, g4.z, b4.z);
                        x[i + 3] = fmaf((x[i + 3] - mean) * rstd, g4.w, b4.w);
                    }
                    store_bf16(col, x);
                }
            } else if (S.epi == TAIL_EPI_ACT) {
                for (int j = 0; j < half / 32; ++j) {
                    load_biased(j, x);
                    const int col = c0 + j * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) x[i] = apply_act(x[i], S.act);
                    if (S.out_f32 && valid) {
                        float* dst = S.out_f32 + static_cast<long long>(grow) * S.ld_f32 + col;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
                    }
                    store_bf16(col, x);
                }
            } else {
                // last hidden layer -> final Linear (GEMV per row) -> softmax (src/multimodal_classifier.py:166-167)
                float acc[kMaxC];
#pragma unroll
                for (int c = 0; c < kMaxC; ++c) acc[c] = 0.0f;
                for (int j = 0; j < half / 32; ++j) {
                    load_biased(j, x);
                    const int col = c0 + j * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) x[i] = apply_act(x[i], S.act);
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c) {
                        if (c < p.C) {
                            const float* w = p.out_w + c * p.h_last + col;
                            float a = acc[c];
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + i));
                                a = fmaf(x[i + 0], w4.x, a);
                                a = fmaf(x[i + 1], w4.y, a);
                                a = fmaf(x[i + 2], w4.z, a);
                                a = fmaf(x[i + 3], w4.w, a);
                            }
                            acc[c] = a;
                        }
                    }
                }
                if (h == 1) {
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c)
                        if (c < p.C) part[row * kMaxC + c] = acc[c];
                }
                named_bar_sync(1, kEpi);
                if (h == 0 && valid) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c)
                        if (c < p.C) {
                            acc[c] += part[row * kMaxC + c] + __ldg(p.out_b + c);
                            mx = fmaxf(mx, acc[c]);
                        }
                    float den = 0.0f;
                    float e[kMaxC];
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c)
                        if (c < p.C) {
                            e[c] = __expf(acc[c] - mx);
                            den += e[c];
                        }
                    const float inv = 1.0f / den;
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c)
                        if (c < p.C) {
                            if (p.logits) p.logits[static_cast<long long>(grow) * p.C + c] = acc[c];
                            if (p.probs) p.probs[static_cast<long long>(grow) * p.C + c] = e[c] * inv;
                        }
                }
            }
            // accumulator drained, A operand of the next stage written: visible to the tensor core's async proxy
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(aready);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// ---------------------------------------------------------------------------------------------- weight preparation
// C[M][N] (row stride ldc) = A[M][K] . B[K][N] in fp32;  cv[i] = sum_k A[i][k] bvec[k] + cin[i]
__global__ void tail_matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int M, int N, int K,
                                       float* __restrict__ Cm, int ldc, const float* __restrict__ bvec,
                                       const float* __restrict__ cin, float* __restrict__ cv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < N) {
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[static_cast<long long>(i) * K + k], Bm[static_cast<long long>(k) * N + j], acc);
        Cm[static_cast<long long>(i) * ldc + j] = acc;
    }
    if (j == 0 && cv) {
        float acc = cin ? cin[i] : 0.0f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[static_cast<long long>(i) * K + k], bvec[k], acc);
        cv[i] = acc;
    }
}

// one block row of w_big: out[i][j] = (P . W)[i][j] for the attended operand, W[i][j] or 0 for the residual operand
__global__ void tail_pack_big_kernel(const float* __restrict__ P, const float* __restrict__ pb,
                                     const float* __restrict__ Watt, const float* __restrict__ batt, int Katt,
                                     const float* __restrict__ Wres, const float* __restrict__ bres, int Kres,
                                     int residual, int F, int att_col0, int res_col0, int ld,
                                     __nv_bfloat16* __restrict__ out, float* __restrict__ bout) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < Katt) {
        float acc = 0.0f;
        for (int k = 0; k < F; ++k) acc = fmaf(P[static_cast<long long>(i) * F + k], Watt[static_cast<long long>(k) * Katt + j], acc);
        out[static_cast<long long>(i) * ld + att_col0 + j] = __float2bfloat16_rn(acc);
    }
    if (j < Kres)
        out[static_cast<long long>(i) * ld + res_col0 + j] =
            __float2bfloat16_rn(residual ? Wres[static_cast<long long>(i) * Kres + j] : 0.0f);
    if (j == 0) {
        float acc = pb[i] + (residual ? bres[i] : 0.0f);
        for (int k = 0; k < F; ++k) acc = fmaf(P[static_cast<long long>(i) * F + k], batt[k], acc);
        bout[i] = acc;
    }
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: %s", what, cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

int pick_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

int weight_map(CUtensorMap* m, const __nv_bfloat16* w, int rows, int K, int bn) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    return encode_tensor_map(m, w, 2, 2, dims, str, box, 128);
}

}  // namespace

bool tail_supported(int F, int img_in, int txt_in, int num_hidden, const int* hdim, int C) {
    if (F % 64 != 0 || F < 64 || F > 512 || img_in % 64 != 0 || txt_in % 64 != 0 || img_in <= 0 || txt_in <= 0)
        return false;
    if (num_hidden < 1 || num_hidden > 3 || C < 1 || C > kMaxC) return false;
    for (int j = 0; j < num_hidden; ++j)
        if (hdim[j] % 64 != 0 || hdim[j] < 64 || hdim[j] > 512) return false;
    return true;
}

int plan_tail(TailLaunch* out, const TailWeights& w, const __nv_bfloat16* img, const __nv_bfloat16* txt, int B) {
    memset(out, 0, sizeof(*out));
    if (!tail_supported(w.F, w.img_in, w.txt_in, w.num_hidden, w.hdim, w.C) || B <= 0) {
        set_last_error("plan_tail: unsupported fusion / head shape (F=%d, img %d, txt %d, %d hidden layers, %d classes)",
                       w.F, w.img_in, w.txt_in, w.num_hidden, w.C);
        return -1;
    }
    TailParams& p = out->p;
    const int F = w.F, Kin = w.img_in + w.txt_in;
    {
        uint64_t dims[2] = {(uint64_t)w.img_in, (uint64_t)B};
        uint64_t str[1] = {(uint64_t)w.img_in * 2};
        uint32_t box[2] = {64, (uint32_t)kRows};
        int rc = encode_tensor_map(&p.img_map, img, 2, 2, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)w.txt_in, (uint64_t)B};
        uint64_t str[1] = {(uint64_t)w.txt_in * 2};
        uint32_t box[2] = {64, (uint32_t)kRows};
        int rc = encode_tensor_map(&p.txt_map, txt, 2, 2, dims, str, box, 128);
        if (rc) return rc;
    }
    p.img_chunks = w.img_in / 64;
    p.B = B;
    p.ln_eps = w.ln_eps;
    const int region[2] = {0, kActChunks / 2};   // the two halves of the activation area, used alternately
    int ns = 0;
    double macs = 0;
    auto add = [&](const __nv_bfloat16* wt, int rows_total, int K, int row0, int n, const float* bias, int a_chunk0,
                   int out_chunk0, int epi, int act) -> int {
        TailStage& S = p.st[ns++];
        S.bn = pick_bn(n);
        int rc = weight_map(&S.w_map, wt, rows_total, K, S.bn);
        if (rc) return rc;
        S.bias = bias;
        S.k_chunks = K / 64;
        S.n = n;
        S.w_row0 = row0;
        S.a_chunk0 = a_chunk0;
        S.out_chunk0 = out_chunk0;
        S.epi = epi;
        S.act = act;
        macs += 1.0 * n * K;
        return 0;
    };
    // pre-LayerNorm rows of both modalities straight from the embeddings, normalised into the concat operand
    int rc = add(w.w_big, 2 * F, Kin, 0, F, w.b_big, -1, 0, TAIL_EPI_LN, ACT_NONE);
    if (rc) return rc;
    p.st[0].ln_g = w.ln_i_g;
    p.st[0].ln_b = w.ln_i_b;
    rc = add(w.w_big, 2 * F, Kin, F, F, w.b_big + F, -1, F / 64, TAIL_EPI_LN, ACT_NONE);
    if (rc) return rc;
    p.st[1].ln_g = w.ln_t_g;
    p.st[1].ln_b = w.ln_t_b;
    // fusion MLP: Linear(2F, F) + ReLU (+ Dropout = identity in eval), Linear(F, F)   (src/fusion_model.py:233-240)
    rc = add(w.w1, F, 2 * F, 0, F, w.b1, 0, region[0], TAIL_EPI_ACT, ACT_RELU);
    if (rc) return rc;
    rc = add(w.w2, F, F, 0, F, w.b2, region[0], region[1], TAIL_EPI_ACT, ACT_NONE);
    if (rc) return rc;
    const int fused_stage = ns - 1;
    (void)fused_stage;
    int in_region = 1, in_dim = F;
    for (int j = 0; j < w.num_hidden; ++j) {
        const bool last = j == w.num_hidden - 1;
        rc = add(w.wh[j], w.hdim[j], in_dim, 0, w.hdim[j], w.bh[j], region[in_region], region[in_region ^ 1],
                 last ? TAIL_EPI_FINAL : TAIL_EPI_ACT, w.head_act);
        if (rc) return rc;
        in_region ^= 1;
        in_dim = w.hdim[j];
    }
    p.num_stages = ns;
    p.out_w = w.out_w;
    p.out_b = w.out_b;
    p.h_last = in_dim;
    p.C = w.C;
    macs += 1.0 * in_dim * w.C;
    out->grid = (B + kRows - 1) / kRows;
    out->flops = 2.0 * B * macs;
    out->bytes = 2.0 * B * Kin + 2.0 * macs + 8.0 * B * w.C;
    return 0;
}

int launch_tail(const TailLaunch* g, float* logits, float* probs, float* fused_f32, int ld_fused, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) {
            set_last_error("tail_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -static_cast<int>(e);
        }
        attr_set = true;
    }
    TailParams p = g->p;
    p.logits = logits;
    p.probs = probs;
    p.st[3].out_f32 = fused_f32;   // stage 3 = fusion.3, the fused embedding (plan_tail)
    p.st[3].ld_f32 = ld_fused;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tail_fused_kernel, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("tail_fused_kernel launch: %s", cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

int pack_tail_big(const float* wip, const float* bip, const float* wtp, const float* btp, const CrossRaw& i2t,
                  const CrossRaw& t2i, int F, int Ii, int Ti, int residual, float* scratch, __nv_bfloat16* w_big,
                  float* b_big, cudaStream_t s) {
    float* P1 = scratch;                       // Wo Wv of image_to_text_attention
    float* P2 = scratch + 1LL * F * F;         // ... of text_to_image_attention
    float* pb1 = P2 + 1LL * F * F;
    float* pb2 = pb1 + F;
    const dim3 gF((F + 127) / 128, F);
    tail_matmul_f32_kernel<<<gF, 128, 0, s>>>(i2t.wo, i2t.wv, F, F, F, P1, F, i2t.bv, i2t.bo, pb1);
    tail_matmul_f32_kernel<<<gF, 128, 0, s>>>(t2i.wo, t2i.wv, F, F, F, P2, F, t2i.bv, t2i.bo, pb2);
    const int ld = Ii + Ti;
    const int kmax = Ii > Ti ? Ii : Ti;
    const dim3 g((kmax + 127) / 128, F);
    // rows [0, F): pre_i = image_proj(img) [residual] + P1 text_proj(txt)
    tail_pack_big_kernel<<<g, 128, 0, s>>>(P1, pb1, wtp, btp, Ti, wip, bip, Ii, residual, F, Ii, 0, ld, w_big, b_big);
    // rows [F, 2F): pre_t = text_proj(txt) [residual] + P2 image_proj(img)
    tail_pack_big_kernel<<<g, 128, 0, s>>>(P2, pb2, wip, bip, Ii, wtp, btp, Ti, residual, F, 0, Ii, ld,
                                           w_big + 1LL * F * ld, b_big + F);
    return check_launch("pack_tail_big");
}

}  // namespace mrd
