// K3: fused masked-softmax self-attention (flash-style, one pass over the keys).
//
// One CTA = one (sample, head, BLOCK_M query rows); each warp owns 16 query rows.  Keys/values are
// streamed in blocks of 64 through a cp.async double buffer in 128B-XOR-swizzled shared memory;
// QK^T and PV run on mma.sync m16n8k16 (bf16 in, fp32 accumulate) with the probabilities kept in
// registers between the two products; softmax statistics are fp32 and online.  The key-padding mask
// is a per-key additive bias ([B,S], 0 / -inf) staged once in shared memory: the [B,1,S,S] mask of
// HF:masking_utils.py:1001-1088 is never built, and 64-key blocks whose keys are all padded are
// skipped outright.  At BERT-base sizes attention is 2.7 % (S=128) to 10 % (S=512) of the forward's
// FLOPs, which is why it stays on the legacy tensor path while the GEMMs/convs use tcgen05.

#include "attention.h"

#include <math.h>
#include <stdint.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace mrd {

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlockN = 64;
constexpr int kMaxS = 512;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                            uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                                  uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows x 64 bf16 tile, global row stride ld (elements) -> swizzled smem (128 B per row)
template <int NT>
__device__ __forceinline__ void load_tile(uint32_t dst, const __nv_bfloat16* base, long long ld,
                                          int row0, int rows, int S, int tid) {
    for (int idx = tid; idx < rows * 8; idx += NT) {
        const int r = idx >> 3, c = idx & 7;
        const int grow = row0 + r;
        const bool ok = grow < S;
        const __nv_bfloat16* src = base + static_cast<long long>(ok ? grow : 0) * ld + c * 8;
        cp_async16(dst + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
    }
}

// seq_off == null: dense layout, sample b owns rows [b*S_max, (b+1)*S_max) and mask_bias is [B,S_max].
// seq_off != null: token-packed layout, sample b owns rows [seq_off[b], seq_off[b+1]) and mask_bias is
// indexed by packed row.
template <int BLOCK_M>
__global__ void __launch_bounds__(BLOCK_M * 2, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ mask_bias,
                 const int* __restrict__ seq_off, int S_max, int heads,
                 __nv_bfloat16* __restrict__ out) {
    constexpr int NT = BLOCK_M * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t q_s = smem_u32(smem);
    const uint32_t k_s = q_s + BLOCK_M * 128;
    const uint32_t v_s = k_s + 2 * kBlockN * 128;
    float* bias_s = reinterpret_cast<float*>(smem + BLOCK_M * 128 + 4 * kBlockN * 128);
    int* kb_list = reinterpret_cast<int*>(bias_s + kMaxS);
    int* nkb_s = kb_list + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int q0 = blockIdx.x * BLOCK_M, h = blockIdx.y, b = blockIdx.z;
    long long off;
    int S;
    if (seq_off) {
        off = __ldg(seq_off + b);
        S = __ldg(seq_off + b + 1) - static_cast<int>(off);
    } else {
        off = static_cast<long long>(b) * S_max;
        S = S_max;
    }
    if (q0 >= S) return;  // uniform for the whole CTA
    const long long ld = 3LL * heads * kHeadDim;
    const __nv_bfloat16* q_base = qkv + off * ld + h * kHeadDim;
    const __nv_bfloat16* k_base = q_base + heads * kHeadDim;
    const __nv_bfloat16* v_base = k_base + heads * kHeadDim;

    // ---- key bias -> smem, list of key blocks that contain at least one attended key
    const int nkb_total = (S + kBlockN - 1) / kBlockN;
    for (int i = tid; i < nkb_total * kBlockN; i += NT)
        bias_s[i] = i < S ? (mask_bias ? __ldg(mask_bias + off + i) : 0.0f) : -INFINITY;
    __syncthreads();
    if (warp == 0) {
        int cnt = 0;
        for (int kb = 0; kb < nkb_total; ++kb) {
            const bool v = bias_s[kb * kBlockN + lane] > -INFINITY ||
                           bias_s[kb * kBlockN + 32 + lane] > -INFINITY;
            if (__any_sync(0xffffffffu, v)) {
                if (lane == 0) kb_list[cnt] = kb;
                ++cnt;
            }
        }
        if (lane == 0) *nkb_s = cnt;
    }
    __syncthreads();
    const int nkb = *nkb_s;

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    uint32_t qf[4][4];

    if (nkb > 0) {
        load_tile<NT>(q_s, q_base, ld, q0, BLOCK_M, S, tid);
        load_tile<NT>(k_s, k_base, ld, kb_list[0] * kBlockN, kBlockN, S, tid);
        load_tile<NT>(v_s, v_base, ld, kb_list[0] * kBlockN, kBlockN, S, tid);
        cp_async_commit();
    }

    for (int it = 0; it < nkb; ++it) {
        const int buf = it & 1;
        if (it + 1 < nkb) {
            const int nb = kb_list[it + 1] * kBlockN;
            load_tile<NT>(k_s + (buf ^ 1) * kBlockN * 128, k_base, ld, nb, kBlockN, S, tid);
            load_tile<NT>(v_s + (buf ^ 1) * kBlockN * 128, v_base, ld, nb, kBlockN, S, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (it == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int r = warp * 16 + (lane & 15);
                const int c = kk * 2 + (lane >> 4);
                ldmatrix_x4(q_s + r * 128 + ((c ^ (r & 7)) << 4), qf[kk][0], qf[kk][1], qf[kk][2],
                            qf[kk][3]);
            }
        }

        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
        const uint32_t kt = k_s + buf * kBlockN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int mi = lane >> 3;
                const int r = np * 16 + (mi >> 1) * 8 + (lane & 7);
                const int c = kk * 2 + (mi & 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(kt + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
                mma_bf16(s[np * 2], qf[kk], b0, b1);
                mma_bf16(s[np * 2 + 1], qf[kk], b2, b3);
            }
        }

        // ---- online softmax (rows g and g+8 of this warp's 16)
        const float* bb = bias_s + kb_list[it] * kBlockN + tq * 2;
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b0 = bb[j * 8], b1 = bb[j * 8 + 1];
            s[j][0] += b0; s[j][1] += b1; s[j][2] += b0; s[j][3] += b1;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mu0 = mx0 == -INFINITY ? 0.0f : mx0;
        const float mu1 = mx1 == -INFINITY ? 0.0f : mx1;
        const float corr0 = exp2f((m0 - mu0) * kLog2e);
        const float corr1 = exp2f((m1 - mu1) * kLog2e);
        m0 = mx0; m1 = mx1;
        l0 *= corr0; l1 *= corr1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr0; o[j][1] *= corr0; o[j][2] *= corr1; o[j][3] *= corr1;
        }
        const float ms0 = mu0 * kLog2e, ms1 = mu1 * kLog2e;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(fmaf(s[j][0], kLog2e, -ms0));
            s[j][1] = exp2f(fmaf(s[j][1], kLog2e, -ms0));
            s[j][2] = exp2f(fmaf(s[j][2], kLog2e, -ms1));
            s[j][3] = exp2f(fmaf(s[j][3], kLog2e, -ms1));
            l0 += s[j][0] + s[j][1];
            l1 += s[j][2] + s[j][3];
        }

        // ---- O += P V
        const uint32_t vt = v_s + buf * kBlockN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t a[4];
            a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                const int mi = lane >> 3;
                const int r = kk * 16 + (mi & 1) * 8 + (lane & 7);
                const int c = dp * 2 + (mi >> 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(vt + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
                mma_bf16(o[dp * 2], a, b0, b1);
                mma_bf16(o[dp * 2 + 1], a, b2, b3);
            }
        }
        __syncthreads();
    }

    // ---- normalise, stage through this warp's own Q rows, 16-byte coalesced stores
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.0f ? 1.0f / l0 : 0.0f;
    const float inv1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;
    uint8_t* q_gen = smem;
    {
        const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<uint32_t*>(q_gen + r0 * 128 + ((j ^ (r0 & 7)) << 4) + tq * 4) =
                pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
            *reinterpret_cast<uint32_t*>(q_gen + r1 * 128 + ((j ^ (r1 & 7)) << 4) + tq * 4) =
                pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
        }
    }
    __syncwarp();
    const long long ldo = static_cast<long long>(heads) * kHeadDim;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = i * 32 + lane;
        const int r = warp * 16 + (idx >> 3), c = idx & 7;
        const int q = q0 + r;
        if (q < S) {
            const uint4 v = *reinterpret_cast<const uint4*>(q_gen + r * 128 + ((c ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(out + (off + q) * ldo + h * kHeadDim + c * 8) = v;
        }
    }
}

template <int BLOCK_M>
int launch(const __nv_bfloat16* qkv, const float* mask_bias, const int* seq_off, int B, int S,
           int heads, __nv_bfloat16* out, cudaStream_t stream) {
    constexpr int SMEM = BLOCK_M * 128 + 4 * kBlockN * 128 + kMaxS * 4 + 64;
    static bool attr_set = false;
    auto kfn = attention_kernel<BLOCK_M>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -static_cast<int>(e);
        }
        attr_set = true;
    }
    dim3 grid((S + BLOCK_M - 1) / BLOCK_M, heads, B);
    kfn<<<grid, BLOCK_M * 2, SMEM, stream>>>(qkv, mask_bias, seq_off, S, heads, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("attention_kernel<%d> launch: %s", BLOCK_M, cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

}  // namespace

int attention_forward(const __nv_bfloat16* qkv, const float* mask_bias, const int* seq_off, int B,
                      int S, int heads, __nv_bfloat16* out, cudaStream_t stream) {
    if (B <= 0 || S <= 0) return 0;
    if (S > kMaxS || heads <= 0 || heads > 65535 || B > 65535) {
        set_last_error("attention_forward: unsupported B=%d S=%d heads=%d (S <= %d)", B, S, heads,
                       kMaxS);
        return -1;
    }
    // 64-row query blocks when sequences are short (or packed to short lengths): a block whose rows
    // all lie beyond the sequence exits immediately; 128-row blocks halve the K/V re-reads otherwise
    if (S > 128 || (S > 64 && seq_off == nullptr))
        return launch<128>(qkv, mask_bias, seq_off, B, S, heads, out, stream);
    return launch<64>(qkv, mask_bias, seq_off, B, S, heads, out, stream);
}

}  // namespace mrd
