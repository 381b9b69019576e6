// Backward of the fused masked-softmax attention (training step).  Reference:
// autograd through BertSelfAttention (HF:models/bert/modeling_bert.py:168-207 ->
// F.scaled_dot_product_attention with dropout_p = attention_probs_dropout_prob in train mode).
//
// One CTA = one (sample, head); sequences of at most 128 tokens, so Q, K, V, dO are single 128x64 tiles
// and the probabilities one 128x128 tile - nothing is streamed and no softmax statistics have to be
// saved by the forward pass: S = Q K^T is recomputed here.
//   phase A (warp = 16 query rows): S, P = softmax(S + bias), dP = dO V^T, delta = rowsum(dO * O),
//            dS = P * (mask/(1-p) * dP - delta); dropped P and dS go to shared memory as bf16
//   phase B (warp = 16 query rows): dQ = dS K
//   phase C (warp = 16 key rows):   dV = Pdrop^T dO,  dK = dS^T Q        (transposed ldmatrix)
// All five products run on mma.sync m16n8k16 (bf16 in, fp32 accumulate).  At BERT-base training sizes
// this is ~3 % of the step; the tcgen05 tiles are spent on the GEMMs.

#include <math.h>
#include <stdint.h>

#include "ptx.cuh"
#include "tma_host.h"
#include "train_kernels.h"

namespace mrd {

namespace {

constexpr int kHeadDim = 64;
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kTile = 128 * 128;            // bytes of one 128x64 bf16 tile
constexpr int kSq = 128 * 256;              // bytes of one 128x128 bf16 tile
constexpr int kSmem = 4 * kTile + 2 * kSq + 128 * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 128 rows x 64 bf16 (row stride ld) -> smem, 128 B per row, 16-byte chunks XOR-swizzled by (row & 7);
// rows >= len are zero-filled
__device__ __forceinline__ void load_tile(uint32_t dst, const __nv_bfloat16* base, long long ld, int len, int tid) {
    const int c = tid & 7;
#pragma unroll
    for (int r = tid >> 3; r < 128; r += kThreads / 8) {
        const bool ok = r < len;
        cp_async16(dst + r * 128 + ((c ^ (r & 7)) << 4), ok ? base + static_cast<long long>(r) * ld + c * 8 : base, ok);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ ctx,
                     const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ mask_bias,
                     const int* __restrict__ seq_off, int S_max, int heads, DropCfg d,
                     __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t q_s = smem_u32(smem), k_s = q_s + kTile, v_s = k_s + kTile, do_s = v_s + kTile;
    const uint32_t p_s = do_s + kTile, ds_s = p_s + kSq;
    uint8_t* p_gen = smem + 4 * kTile;
    uint8_t* ds_gen = p_gen + kSq;
    uint8_t* do_gen = smem + 3 * kTile;
    float* bias_s = reinterpret_cast<float*>(smem + 4 * kTile + 2 * kSq);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
    long long off;
    int len;
    if (seq_off) {
        off = __ldg(seq_off + b);
        len = __ldg(seq_off + b + 1) - static_cast<int>(off);
    } else {
        off = static_cast<long long>(b) * S_max;
        len = S_max;
    }
    if (len <= 0) return;
    if (len > 128) len = 128;
    const long long ld = 3LL * heads * kHeadDim, ldo = static_cast<long long>(heads) * kHeadDim;
    const __nv_bfloat16* qrow = qkv + off * ld + h * kHeadDim;
    load_tile(q_s, qrow, ld, len, tid);
    load_tile(k_s, qrow + heads * kHeadDim, ld, len, tid);
    load_tile(v_s, qrow + 2 * heads * kHeadDim, ld, len, tid);
    load_tile(do_s, dctx + off * ldo + h * kHeadDim, ldo, len, tid);
    cp_async_commit();
    if (tid < 128) bias_s[tid] = tid < len ? (mask_bias ? __ldg(mask_bias + off + tid) : 0.0f) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();

    const int n16 = (len + 15) >> 4;  // 16-row groups that hold live queries / keys
    const int mi = lane >> 3, x7 = lane & 7;
    // B operand straight from [n][k] rows (K^T, V^T products) / transposed from [k][n] rows (K, Q, dO products)
    const uint32_t kn_row = static_cast<uint32_t>(((mi >> 1) * 8 + x7) * 128);
    const uint32_t kt_row = static_cast<uint32_t>(((mi & 1) * 8 + x7) * 128);
    uint32_t kn_col[4], kt_col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        kn_col[i] = static_cast<uint32_t>(((i * 2 + (mi & 1)) ^ x7) << 4);
        kt_col[i] = static_cast<uint32_t>(((i * 2 + (mi >> 1)) ^ x7) << 4);
    }
    const int r_lo = warp * 16 + g, r_hi = r_lo + 8;   // the two rows whose accumulators this thread holds

    // ------------------------------------------------------------------ phase A
    if (warp < n16) {
        uint32_t qf[4][4], dof[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int r = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
            const uint32_t o = r * 128 + ((c ^ (r & 7)) << 4);
            ldsm(q_s + o, qf[kk]);
            ldsm(do_s + o, dof[kk]);
        }
        float sacc[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                if (np >= n16) break;
                uint32_t bb[4];
                ldsm(k_s + np * 2048 + kn_row + kn_col[kk], bb);
                mma_bf16(sacc[np * 2], qf[kk], bb[0], bb[1]);
                mma_bf16(sacc[np * 2 + 1], qf[kk], bb[2], bb[3]);
            }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float b0 = bias_s[j * 8 + tq * 2], b1 = bias_s[j * 8 + tq * 2 + 1];
            sacc[j][0] += b0; sacc[j][1] += b1; sacc[j][2] += b0; sacc[j][3] += b1;
            mx0 = fmaxf(mx0, fmaxf(sacc[j][0], sacc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float ms0 = (mx0 == -INFINITY ? 0.0f : mx0) * kLog2e, ms1 = (mx1 == -INFINITY ? 0.0f : mx1) * kLog2e;
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            sacc[j][0] = fast_exp2(fmaf(sacc[j][0], kLog2e, -ms0));
            sacc[j][1] = fast_exp2(fmaf(sacc[j][1], kLog2e, -ms0));
            sacc[j][2] = fast_exp2(fmaf(sacc[j][2], kLog2e, -ms1));
            sacc[j][3] = fast_exp2(fmaf(sacc[j][3], kLog2e, -ms1));
            l0 += sacc[j][0] + sacc[j][1];
            l1 += sacc[j][2] + sacc[j][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, inv1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;

        // delta = rowsum(dO * O): the quad of a row splits the 64 dims, 16 each
        float dl0 = 0.0f, dl1 = 0.0f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = half ? r_hi : r_lo;
            float acc = 0.0f;
            if (r < len) {
                const __nv_bfloat16* op = ctx + (off + r) * ldo + h * kHeadDim + tq * 16;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = tq * 2 + cc;
                    const uint4 dv = *reinterpret_cast<const uint4*>(do_gen + r * 128 + ((c ^ (r & 7)) << 4));
                    const uint4 ov = __ldg(reinterpret_cast<const uint4*>(op) + cc);
                    const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = unpack_bf16(dw[e]), bq = unpack_bf16(ow[e]);
                        acc = fmaf(a.x, bq.x, acc);
                        acc = fmaf(a.y, bq.y, acc);
                    }
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (half) dl1 = acc; else dl0 = acc;
        }

        const unsigned long long idx_bh = static_cast<unsigned long long>(b * heads + h) * S_max;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float dp[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.0f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (half * 4 + np >= n16) break;
                    uint32_t bb[4];
                    ldsm(v_s + (half * 4 + np) * 2048 + kn_row + kn_col[kk], bb);
                    mma_bf16(dp[np * 2], dof[kk], bb[0], bb[1]);
                    mma_bf16(dp[np * 2 + 1], dof[kk], bb[2], bb[3]);
                }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = half * 8 + jj;
                const int key = j * 8 + tq * 2;
                float pd[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = (e < 2) ? r_lo : r_hi;
                    const float p = sacc[j][e] * ((e < 2) ? inv0 : inv1);
                    const bool keep = drop_keep(d, (idx_bh + r) * S_max + key + (e & 1));
                    pd[e] = keep ? p * d.scale : 0.0f;
                    const float dpe = keep ? dp[jj][e] * d.scale : 0.0f;
                    ds[e] = p * (dpe - ((e < 2) ? dl0 : dl1));
                }
                const uint32_t o_lo = r_lo * 256 + ((j ^ (r_lo & 7)) << 4) + tq * 4;
                const uint32_t o_hi = r_hi * 256 + ((j ^ (r_hi & 7)) << 4) + tq * 4;
                *reinterpret_cast<uint32_t*>(p_gen + o_lo) = pack_bf16(pd[0], pd[1]);
                *reinterpret_cast<uint32_t*>(p_gen + o_hi) = pack_bf16(pd[2], pd[3]);
                *reinterpret_cast<uint32_t*>(ds_gen + o_lo) = pack_bf16(ds[0], ds[1]);
                *reinterpret_cast<uint32_t*>(ds_gen + o_hi) = pack_bf16(ds[2], ds[3]);
            }
        }
    }
    __syncthreads();
    if (warp >= n16) return;   // neither live queries nor live keys in this warp's 16 rows

    // ------------------------------------------------------------------ phase B: dQ = dS K
    {
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
        for (int kk = 0; kk < n16; ++kk) {
            uint32_t a[4];
            const int r = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
            ldsm(ds_s + r * 256 + ((c ^ (r & 7)) << 4), a);
#pragma unroll
            for (int dpi = 0; dpi < 4; ++dpi) {
                uint32_t bb[4];
                ldsm_t(k_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                mma_bf16(acc[dpi * 2], a, bb[0], bb[1]);
                mma_bf16(acc[dpi * 2 + 1], a, bb[2], bb[3]);
            }
        }
        __nv_bfloat16* dq = dqkv + off * ld + h * kHeadDim;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = j * 8 + tq * 2;
            if (r_lo < len) *reinterpret_cast<uint32_t*>(dq + r_lo * ld + col) = pack_bf16(acc[j][0], acc[j][1]);
            if (r_hi < len) *reinterpret_cast<uint32_t*>(dq + r_hi * ld + col) = pack_bf16(acc[j][2], acc[j][3]);
        }
    }

    // ------------------------------------------------------------------ phase C: dV = Pdrop^T dO, dK = dS^T Q
    {
        float dv[8][4], dk[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.0f;
            dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.0f;
        }
        for (int kk = 0; kk < n16; ++kk) {   // 16 queries per step
            const int qr = kk * 16 + (mi >> 1) * 8 + x7;
            const uint32_t o = qr * 256 + (((2 * warp + (mi & 1)) ^ x7) << 4);
            uint32_t ap[4], as[4];
            ldsm_t(p_s + o, ap);
            ldsm_t(ds_s + o, as);
#pragma unroll
            for (int dpi = 0; dpi < 4; ++dpi) {
                uint32_t bb[4];
                ldsm_t(do_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                mma_bf16(dv[dpi * 2], ap, bb[0], bb[1]);
                mma_bf16(dv[dpi * 2 + 1], ap, bb[2], bb[3]);
                ldsm_t(q_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                mma_bf16(dk[dpi * 2], as, bb[0], bb[1]);
                mma_bf16(dk[dpi * 2 + 1], as, bb[2], bb[3]);
            }
        }
        __nv_bfloat16* dkp = dqkv + off * ld + heads * kHeadDim + h * kHeadDim;
        __nv_bfloat16* dvp = dkp + heads * kHeadDim;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = j * 8 + tq * 2;
            if (r_lo < len) {
                *reinterpret_cast<uint32_t*>(dkp + r_lo * ld + col) = pack_bf16(dk[j][0], dk[j][1]);
                *reinterpret_cast<uint32_t*>(dvp + r_lo * ld + col) = pack_bf16(dv[j][0], dv[j][1]);
            }
            if (r_hi < len) {
                *reinterpret_cast<uint32_t*>(dkp + r_hi * ld + col) = pack_bf16(dk[j][2], dk[j][3]);
                *reinterpret_cast<uint32_t*>(dvp + r_hi * ld + col) = pack_bf16(dv[j][2], dv[j][3]);
            }
        }
    }
}

// ====================================================================== sequences of 129..512 tokens
// One CTA = one (sample, head, block of 128 queries).  Pass 1 walks the key blocks once to get the softmax
// statistics of its query rows (online max / sum; nothing was saved by the forward); pass 2 walks them
// again doing the work of the single-tile kernel per (query block, key block) pair: dQ accumulates in
// registers across the key blocks, dK / dV of a key block are added into an fp32 accumulator in global
// memory (several query blocks contribute) that the caller converts to bf16 afterwards.
constexpr int kSmemTiled = kSmem + 512 * 4;

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ ctx,
                           const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ mask_bias,
                           const int* __restrict__ seq_off, int S_max, int heads, DropCfg d,
                           __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dkv_acc) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t q_s = smem_u32(smem), k_s = q_s + kTile, v_s = k_s + kTile, do_s = v_s + kTile;
    const uint32_t p_s = do_s + kTile, ds_s = p_s + kSq;
    uint8_t* p_gen = smem + 4 * kTile;
    uint8_t* ds_gen = p_gen + kSq;
    uint8_t* do_gen = smem + 3 * kTile;
    float* bias_s = reinterpret_cast<float*>(smem + 4 * kTile + 2 * kSq);   // 512 keys

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
    const int qb = blockIdx.y;
    long long off;
    int len;
    if (seq_off) {
        off = __ldg(seq_off + b);
        len = __ldg(seq_off + b + 1) - static_cast<int>(off);
    } else {
        off = static_cast<long long>(b) * S_max;
        len = S_max;
    }
    if (len > 512) len = 512;
    const int q_base = qb * 128;
    if (q_base >= len) return;   // uniform
    const int q_live = min(128, len - q_base);
    const int nkb = (len + 127) >> 7;
    const long long ld = 3LL * heads * kHeadDim, ldo = static_cast<long long>(heads) * kHeadDim;
    const __nv_bfloat16* qrow = qkv + off * ld + h * kHeadDim;
    load_tile(q_s, qrow + static_cast<long long>(q_base) * ld, ld, q_live, tid);
    load_tile(do_s, dctx + (off + q_base) * ldo + h * kHeadDim, ldo, q_live, tid);
    cp_async_commit();
    for (int i = tid; i < 512; i += kThreads)
        bias_s[i] = i < len ? (mask_bias ? __ldg(mask_bias + off + i) : 0.0f) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();

    const int n16q = (q_live + 15) >> 4;
    const int mi = lane >> 3, x7 = lane & 7;
    const uint32_t kn_row = static_cast<uint32_t>(((mi >> 1) * 8 + x7) * 128);
    const uint32_t kt_row = static_cast<uint32_t>(((mi & 1) * 8 + x7) * 128);
    uint32_t kn_col[4], kt_col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        kn_col[i] = static_cast<uint32_t>(((i * 2 + (mi & 1)) ^ x7) << 4);
        kt_col[i] = static_cast<uint32_t>(((i * 2 + (mi >> 1)) ^ x7) << 4);
    }
    const int r_lo = warp * 16 + g, r_hi = r_lo + 8;   // rows within the query block
    const bool q_warp = warp < n16q;

    uint32_t qf[4][4], dof[4][4];
    if (q_warp) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int r = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
            const uint32_t o = r * 128 + ((c ^ (r & 7)) << 4);
            ldsm(q_s + o, qf[kk]);
            ldsm(do_s + o, dof[kk]);
        }
    }
    // delta = rowsum(dO * O)
    float dl0 = 0.0f, dl1 = 0.0f;
    if (q_warp) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = half ? r_hi : r_lo;
            float acc = 0.0f;
            if (r < q_live) {
                const __nv_bfloat16* op = ctx + (off + q_base + r) * ldo + h * kHeadDim + tq * 16;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = tq * 2 + cc;
                    const uint4 dv = *reinterpret_cast<const uint4*>(do_gen + r * 128 + ((c ^ (r & 7)) << 4));
                    const uint4 ov = __ldg(reinterpret_cast<const uint4*>(op) + cc);
                    const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = unpack_bf16(dw[e]), bq = unpack_bf16(ow[e]);
                        acc = fmaf(a.x, bq.x, acc);
                        acc = fmaf(a.y, bq.y, acc);
                    }
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (half) dl1 = acc; else dl0 = acc;
        }
    }

    // S block of this warp's 16 query rows against key block j (bias added)
    auto scores = [&](int j, int n16k, float (&sacc)[16][4]) {
#pragma unroll
        for (int t = 0; t < 16; ++t) sacc[t][0] = sacc[t][1] = sacc[t][2] = sacc[t][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                if (np >= n16k) break;
                uint32_t bb[4];
                ldsm(k_s + np * 2048 + kn_row + kn_col[kk], bb);
                mma_bf16(sacc[np * 2], qf[kk], bb[0], bb[1]);
                mma_bf16(sacc[np * 2 + 1], qf[kk], bb[2], bb[3]);
            }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const float b0 = bias_s[j * 128 + t * 8 + tq * 2], b1 = bias_s[j * 128 + t * 8 + tq * 2 + 1];
            sacc[t][0] += b0; sacc[t][1] += b1; sacc[t][2] += b0; sacc[t][3] += b1;
        }
    };

    // ------------------------------------------------------------------ pass 1: row max / sum over all keys
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    for (int j = 0; j < nkb; ++j) {
        const int k_live = min(128, len - j * 128);
        __syncthreads();   // previous K tile fully consumed
        load_tile(k_s, qrow + heads * kHeadDim + static_cast<long long>(j) * 128 * ld, ld, k_live, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (q_warp) {
            float sacc[16][4];
            scores(j, (k_live + 15) >> 4, sacc);
            float mx0 = m0, mx1 = m1;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                mx0 = fmaxf(mx0, fmaxf(sacc[t][0], sacc[t][1]));
                mx1 = fmaxf(mx1, fmaxf(sacc[t][2], sacc[t][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float u0 = mx0 == -INFINITY ? 0.0f : mx0, u1 = mx1 == -INFINITY ? 0.0f : mx1;
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                a0 += fast_exp2((sacc[t][0] - u0) * kLog2e) + fast_exp2((sacc[t][1] - u0) * kLog2e);
                a1 += fast_exp2((sacc[t][2] - u1) * kLog2e) + fast_exp2((sacc[t][3] - u1) * kLog2e);
            }
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            l0 = l0 * fast_exp2((m0 - u0) * kLog2e) + a0;
            l1 = l1 * fast_exp2((m1 - u1) * kLog2e) + a1;
            m0 = mx0; m1 = mx1;
        }
    }
    const float ms0 = (m0 == -INFINITY ? 0.0f : m0) * kLog2e, ms1 = (m1 == -INFINITY ? 0.0f : m1) * kLog2e;
    const float inv0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, inv1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;

    // ------------------------------------------------------------------ pass 2
    float dq[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.0f;
    const unsigned long long idx_bh = static_cast<unsigned long long>(b * heads + h) * S_max;
    for (int j = 0; j < nkb; ++j) {
        const int k_live = min(128, len - j * 128);
        const int n16k = (k_live + 15) >> 4;
        __syncthreads();   // K / V / P / dS of the previous key block fully consumed
        load_tile(k_s, qrow + heads * kHeadDim + static_cast<long long>(j) * 128 * ld, ld, k_live, tid);
        load_tile(v_s, qrow + 2 * heads * kHeadDim + static_cast<long long>(j) * 128 * ld, ld, k_live, tid);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (q_warp) {
            float sacc[16][4];
            scores(j, n16k, sacc);
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                sacc[t][0] = fast_exp2(fmaf(sacc[t][0], kLog2e, -ms0)) * inv0;
                sacc[t][1] = fast_exp2(fmaf(sacc[t][1], kLog2e, -ms0)) * inv0;
                sacc[t][2] = fast_exp2(fmaf(sacc[t][2], kLog2e, -ms1)) * inv1;
                sacc[t][3] = fast_exp2(fmaf(sacc[t][3], kLog2e, -ms1)) * inv1;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float dp[8][4];
#pragma unroll
                for (int t = 0; t < 8; ++t) dp[t][0] = dp[t][1] = dp[t][2] = dp[t][3] = 0.0f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        if (half * 4 + np >= n16k) break;
                        uint32_t bb[4];
                        ldsm(v_s + (half * 4 + np) * 2048 + kn_row + kn_col[kk], bb);
                        mma_bf16(dp[np * 2], dof[kk], bb[0], bb[1]);
                        mma_bf16(dp[np * 2 + 1], dof[kk], bb[2], bb[3]);
                    }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int t = half * 8 + jj;
                    const int key = j * 128 + t * 8 + tq * 2;
                    float pd[4], ds[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = q_base + ((e < 2) ? r_lo : r_hi);
                        const float p = sacc[t][e];
                        const bool keep = drop_keep(d, (idx_bh + r) * S_max + key + (e & 1));
                        pd[e] = keep ? p * d.scale : 0.0f;
                        const float dpe = keep ? dp[jj][e] * d.scale : 0.0f;
                        ds[e] = p * (dpe - ((e < 2) ? dl0 : dl1));
                    }
                    const uint32_t o_lo = r_lo * 256 + ((t ^ (r_lo & 7)) << 4) + tq * 4;
                    const uint32_t o_hi = r_hi * 256 + ((t ^ (r_hi & 7)) << 4) + tq * 4;
                    *reinterpret_cast<uint32_t*>(p_gen + o_lo) = pack_bf16(pd[0], pd[1]);
                    *reinterpret_cast<uint32_t*>(p_gen + o_hi) = pack_bf16(pd[2], pd[3]);
                    *reinterpret_cast<uint32_t*>(ds_gen + o_lo) = pack_bf16(ds[0], ds[1]);
                    *reinterpret_cast<uint32_t*>(ds_gen + o_hi) = pack_bf16(ds[2], ds[3]);
                }
            }
        }
        __syncthreads();
        if (q_warp) {   // dQ += dS K_j
            for (int kk = 0; kk < n16k; ++kk) {
                uint32_t a[4];
                const int r = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
                ldsm(ds_s + r * 256 + ((c ^ (r & 7)) << 4), a);
#pragma unroll
                for (int dpi = 0; dpi < 4; ++dpi) {
                    uint32_t bb[4];
                    ldsm_t(k_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                    mma_bf16(dq[dpi * 2], a, bb[0], bb[1]);
                    mma_bf16(dq[dpi * 2 + 1], a, bb[2], bb[3]);
                }
            }
        }
        if (warp < n16k) {   // dV_j += Pdrop^T dO_i, dK_j += dS^T Q_i  (this warp's 16 key rows)
            float dv[8][4], dk[8][4];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                dv[t][0] = dv[t][1] = dv[t][2] = dv[t][3] = 0.0f;
                dk[t][0] = dk[t][1] = dk[t][2] = dk[t][3] = 0.0f;
            }
            for (int kk = 0; kk < n16q; ++kk) {
                const int qr = kk * 16 + (mi >> 1) * 8 + x7;
                const uint32_t o = qr * 256 + (((2 * warp + (mi & 1)) ^ x7) << 4);
                uint32_t ap[4], as[4];
                ldsm_t(p_s + o, ap);
                ldsm_t(ds_s + o, as);
#pragma unroll
                for (int dpi = 0; dpi < 4; ++dpi) {
                    uint32_t bb[4];
                    ldsm_t(do_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                    mma_bf16(dv[dpi * 2], ap, bb[0], bb[1]);
                    mma_bf16(dv[dpi * 2 + 1], ap, bb[2], bb[3]);
                    ldsm_t(q_s + kk * 2048 + kt_row + kt_col[dpi], bb);
                    mma_bf16(dk[dpi * 2], as, bb[0], bb[1]);
                    mma_bf16(dk[dpi * 2 + 1], as, bb[2], bb[3]);
                }
            }
            const long long ldacc = 2LL * heads * kHeadDim;
            float* kacc = dkv_acc + (off + j * 128) * ldacc + h * kHeadDim;
            float* vacc = kacc + heads * kHeadDim;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int col = t * 8 + tq * 2;
                if (r_lo < k_live) {
                    atomicAdd(reinterpret_cast<float2*>(kacc + r_lo * ldacc + col), make_float2(dk[t][0], dk[t][1]));
                    atomicAdd(reinterpret_cast<float2*>(vacc + r_lo * ldacc + col), make_float2(dv[t][0], dv[t][1]));
                }
                if (r_hi < k_live) {
                    atomicAdd(reinterpret_cast<float2*>(kacc + r_hi * ldacc + col), make_float2(dk[t][2], dk[t][3]));
                    atomicAdd(reinterpret_cast<float2*>(vacc + r_hi * ldacc + col), make_float2(dv[t][2], dv[t][3]));
                }
            }
        }
    }
    if (q_warp) {
        __nv_bfloat16* dqp = dqkv + (off + q_base) * ld + h * kHeadDim;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int col = t * 8 + tq * 2;
            if (r_lo < q_live) *reinterpret_cast<uint32_t*>(dqp + r_lo * ld + col) = pack_bf16(dq[t][0], dq[t][1]);
            if (r_hi < q_live) *reinterpret_cast<uint32_t*>(dqp + r_hi * ld + col) = pack_bf16(dq[t][2], dq[t][3]);
        }
    }
}

}  // namespace

int attention_backward(const __nv_bfloat16* qkv, const __nv_bfloat16* ctx, const __nv_bfloat16* dctx,
                       const float* mask_bias, const int* seq_off, int B, int S, int heads, DropCfg d,
                       __nv_bfloat16* dqkv, cudaStream_t s, float* dkv_acc) {
    if (B <= 0 || S <= 0) return 0;
    if (S > 512) {
        set_last_error("attention_backward: sequences longer than 512 tokens are not supported (S=%d)", S);
        return -1;
    }
    if (S > 128 && !dkv_acc) {
        set_last_error("attention_backward: S > 128 needs the fp32 dK/dV accumulator");
        return -1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTiled);
        if (e != cudaSuccess) {
            set_last_error("attention_backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -static_cast<int>(e);
        }
        attr_set = true;
    }
    if (S <= 128) {
        attention_bwd_kernel<<<B * heads, kThreads, kSmem, s>>>(qkv, ctx, dctx, mask_bias, seq_off, S, heads, d, dqkv);
    } else {
        dim3 grid(B * heads, (S + 127) / 128);
        attention_bwd_tiled_kernel<<<grid, kThreads, kSmemTiled, s>>>(qkv, ctx, dctx, mask_bias, seq_off, S, heads, d,
                                                                      dqkv, dkv_acc);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention_bwd kernel launch: %s", cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

}  // namespace mrd
