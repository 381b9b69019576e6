// Memory-bound kernels of the forward (see elementwise.h).  Every kernel moves 16 bytes per thread
// per access, keeps warps on consecutive addresses and sizes its grid from the element count; the
// roofline that bounds them is HBM (or L2 when the micro-batch is L2 resident).

#include "elementwise.h"

#include <math.h>
#include <stdint.h>

#include <mrd_b200.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace mrd {

namespace {

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s launch: %s", what, cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

inline unsigned blocks_for(long long n, int per_block) {
    return static_cast<unsigned>((n + per_block - 1) / per_block);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]);
    u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]);
    u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ------------------------------------------------------------------ image repack
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}

// One thread = four consecutive padded pixels of one padded row (Wp % 4 == 0): 32 contiguous output bytes, the three
// colour planes read as coalesced 4-byte loads.  32-bit index arithmetic (blockIdx.y = image): the 64-bit divisions of
// the one-pixel-per-thread version held this HBM-bound pass at 3.0 TB/s.
template <typename T>
__global__ void repack_images_kernel(const T* __restrict__ x, int H, int W, uint4* __restrict__ xpad) {
    const int Hp = H + 6, Wq = (W + 8) >> 2;                  // quads per padded row
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Hp * Wq) return;
    const int hp = q / Wq, wq = q - hp * Wq;
    const int n = blockIdx.y;
    const int h = hp - 3;
    const int plane = H * W;
    const T* img = x + static_cast<long long>(n) * 3 * plane;
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int w = wq * 4 + j - 3;
        uint32_t lo = 0u, hi = 0u;
        if (h >= 0 && h < H && w >= 0 && w < W) {
            const T* p = img + h * W + w;
            lo = pack_bf16(ld_as_float<T>(p), ld_as_float<T>(p + plane));
            hi = pack_bf16(ld_as_float<T>(p + 2 * plane), 0.0f);
        }
        o[2 * j] = lo;
        o[2 * j + 1] = hi;
    }
    uint4* dst = xpad + (static_cast<long long>(n) * Hp * Wq + q) * 2;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ------------------------------------------------------------------ pooling
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8,
                                    uint4* __restrict__ y) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = static_cast<long long>(N) * Ho * Wo * C8;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C8);
    long long r = i / C8;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
    __nv_bfloat162 m0 = ninf, m1 = ninf, m2 = ninf, m3 = ninf;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
        const int h = 2 * ho + dh;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
            const int w = 2 * wo + dw;
            if (w < 0 || w >= W) continue;
            const uint4 v = __ldg(x + ((static_cast<long long>(n) * H + h) * W + w) * C8 + c);
            m0 = __hmax2(m0, *reinterpret_cast<const __nv_bfloat162*>(&v.x));
            m1 = __hmax2(m1, *reinterpret_cast<const __nv_bfloat162*>(&v.y));
            m2 = __hmax2(m2, *reinterpret_cast<const __nv_bfloat162*>(&v.z));
            m3 = __hmax2(m3, *reinterpret_cast<const __nv_bfloat162*>(&v.w));
        }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m0);
    o.y = *reinterpret_cast<uint32_t*>(&m1);
    o.z = *reinterpret_cast<uint32_t*>(&m2);
    o.w = *reinterpret_cast<uint32_t*>(&m3);
    y[i] = o;
}

__global__ void global_avgpool_kernel(const uint4* __restrict__ x, int N, int HW, int C8,
                                      uint4* __restrict__ y_bf16, float* __restrict__ y_f32) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(N) * C8) return;
    const int c = static_cast<int>(i % C8);
    const int n = static_cast<int>(i / C8);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint4* p = x + static_cast<long long>(n) * HW * C8 + c;
#pragma unroll 7
    for (int k = 0; k < HW; ++k) {
        float f[8];
        unpack8(__ldg(p + static_cast<long long>(k) * C8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    const float inv = 1.0f / static_cast<float>(HW);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    if (y_bf16) y_bf16[i] = pack8(acc);
    if (y_f32) {
        float4* o = reinterpret_cast<float4*>(y_f32 + i * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ------------------------------------------------------------------ LayerNorm (+ residual)
// One warp per row; lane l owns the 8-element chunks (c*32 + l), c < NCHUNK, so every warp-wide
// access is a contiguous 512-byte segment.  Two-pass statistics in registers (fp32).
template <int NCHUNK>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                 const __nv_bfloat16* __restrict__ res, long long ldr,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 int rows, __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ y32,
                 long long ldy32, const int* __restrict__ dyn_rows) {
    constexpr int WIDTH = NCHUNK * 256;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (dyn_rows) rows = min(rows, __ldg(dyn_rows));
    if (row >= rows) return;
    float v[NCHUNK][8];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (c * 32 + lane) * 8;
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + col)), v[c]);
        if (res) {
            float r[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(res + row * ldr + col)), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[c][j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[c][j];
    }
    const float mean = warp_sum(sum) * (1.0f / WIDTH);
    float sq = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[c][j] - mean;
            sq += d * d;
        }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / WIDTH) + eps);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (c * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
        float o[8];
        o[0] = (v[c][0] - mean) * rstd * g0.x + b0.x;
        o[1] = (v[c][1] - mean) * rstd * g0.y + b0.y;
        o[2] = (v[c][2] - mean) * rstd * g0.z + b0.z;
        o[3] = (v[c][3] - mean) * rstd * g0.w + b0.w;
        o[4] = (v[c][4] - mean) * rstd * g1.x + b1.x;
        o[5] = (v[c][5] - mean) * rstd * g1.y + b1.y;
        o[6] = (v[c][6] - mean) * rstd * g1.z + b1.z;
        o[7] = (v[c][7] - mean) * rstd * g1.w + b1.w;
        if (y) *reinterpret_cast<uint4*>(y + row * ldy + col) = pack8(o);
        if (y32) {
            float4* p = reinterpret_cast<float4*>(y32 + row * ldy32 + col);
            p[0] = make_float4(o[0], o[1], o[2], o[3]);
            p[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// ------------------------------------------------------------------ BERT embeddings + LayerNorm
__global__ void __launch_bounds__(256)
bert_embed_ln_kernel(const long long* __restrict__ ids, int tokens, int S,
                     const __nv_bfloat16* __restrict__ word, const float* __restrict__ pos_type,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     int vocab, __nv_bfloat16* __restrict__ y, const int* __restrict__ row_tok,
                     const int* __restrict__ dyn_rows) {
    constexpr int WIDTH = 768, NCHUNK = 3;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (dyn_rows) tokens = min(tokens, __ldg(dyn_rows));
    if (row >= tokens) return;
    const long long tok = row_tok ? __ldg(row_tok + row) : row;
    long long id = __ldg(ids + tok);
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int pos = static_cast<int>(tok % S);
    float v[NCHUNK][8];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (c * 32 + lane) * 8;
        unpack8(__ldg(reinterpret_cast<const uint4*>(word + id * WIDTH + col)), v[c]);
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos_type + pos * WIDTH + col));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos_type + pos * WIDTH + col + 4));
        v[c][0] += p0.x; v[c][1] += p0.y; v[c][2] += p0.z; v[c][3] += p0.w;
        v[c][4] += p1.x; v[c][5] += p1.y; v[c][6] += p1.z; v[c][7] += p1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[c][j];
    }
    const float mean = warp_sum(sum) * (1.0f / WIDTH);
    float sq = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[c][j] - mean;
            sq += d * d;
        }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / WIDTH) + eps);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (c * 32 + lane) * 8;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = (v[c][j] - mean) * rstd * __ldg(gamma + col + j) + __ldg(beta + col + j);
        *reinterpret_cast<uint4*>(y + row * WIDTH + col) = pack8(o);
    }
}

// ------------------------------------------------------------------ token packing
__device__ __forceinline__ bool mask_nonzero(const void* mask, int dtype, long long i) {
    switch (dtype) {
        case MRD_DT_I64: return static_cast<const long long*>(mask)[i] != 0;
        case MRD_DT_I32: return static_cast<const int*>(mask)[i] != 0;
        case MRD_DT_F32: return static_cast<const float*>(mask)[i] != 0.0f;
        case MRD_DT_BF16: return __bfloat162float(static_cast<const __nv_bfloat16*>(mask)[i]) != 0.0f;
        default: return static_cast<const unsigned char*>(mask)[i] != 0;
    }
}

// one warp per sequence: number of kept tokens
__global__ void compact_count_kernel(const void* __restrict__ mask, int dtype, int B, int S,
                                     int keep_all, int* __restrict__ counts) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int n = 0;
    if (keep_all || mask == nullptr) {
        n = S;
    } else {
        for (int j0 = 0; j0 < S; j0 += 32) {
            const int j = j0 + lane;
            const bool keep = j < S && (j == 0 || mask_nonzero(mask, dtype, static_cast<long long>(b) * S + j));
            n += __popc(__ballot_sync(0xffffffffu, keep));
        }
    }
    if (lane == 0) counts[b] = n;
}

// single block: exclusive scan of counts -> seq_off[0..B], n_rows
__global__ void compact_scan_kernel(const int* __restrict__ counts, int B, int* __restrict__ seq_off,
                                    int* __restrict__ n_rows) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < B; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < B ? counts[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_tot[lane] = w;  // inclusive totals per warp
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = x + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
        if (i < B) seq_off[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        seq_off[B] = carry_s;
        n_rows[0] = carry_s;
    }
}

// one warp per sequence: packed row -> token index and key bias
__global__ void compact_fill_kernel(const void* __restrict__ mask, int dtype, int B, int S,
                                    int keep_all, const int* __restrict__ seq_off,
                                    int* __restrict__ row_tok, float* __restrict__ row_bias) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int r = seq_off[b];
    for (int j0 = 0; j0 < S; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < S && (mask == nullptr || mask_nonzero(mask, dtype, static_cast<long long>(b) * S + j));
        const bool keep = j < S && (keep_all || mask == nullptr || valid || j == 0);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int dst = r + __popc(bal & ((1u << lane) - 1u));
            row_tok[dst] = b * S + j;
            row_bias[dst] = valid ? 0.0f : -INFINITY;
        }
        r += __popc(bal);
    }
}

__global__ void gather_cls_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ seq_off,
                                  int B, int width8, uint4* __restrict__ y_bf16,
                                  float* __restrict__ y_f32) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * width8) return;
    const int c = static_cast<int>(i % width8);
    const int b = static_cast<int>(i / width8);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(seq_off[b]) * width8 * 8) + c);
    if (y_bf16) y_bf16[i] = v;
    if (y_f32) {
        float f[8];
        unpack8(v, f);
        float4* o = reinterpret_cast<float4*>(y_f32 + i * 8);
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
}

// ------------------------------------------------------------------ mask -> additive bias
__global__ void mask_to_bias_kernel(const void* __restrict__ mask, int dtype, long long n,
                                    float* __restrict__ bias) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool valid;
    switch (dtype) {
        case MRD_DT_I64: valid = static_cast<const long long*>(mask)[i] != 0; break;
        case MRD_DT_I32: valid = static_cast<const int*>(mask)[i] != 0; break;
        case MRD_DT_F32: valid = static_cast<const float*>(mask)[i] != 0.0f; break;
        case MRD_DT_BF16:
            valid = __bfloat162float(static_cast<const __nv_bfloat16*>(mask)[i]) != 0.0f;
            break;
        default: valid = static_cast<const unsigned char*>(mask)[i] != 0; break;
    }
    bias[i] = valid ? 0.0f : -INFINITY;
}

// ------------------------------------------------------------------ classifier tail
__global__ void __launch_bounds__(256)
head_logits_softmax_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                           const float* __restrict__ W, const float* __restrict__ b, int B, int K,
                           int C, float* __restrict__ logits, float* __restrict__ probs) {
    extern __shared__ float w_s[];  // [C][K]
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) w_s[i] = __ldg(W + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    const int per = K >> 5;  // elements per lane, <= 32
    for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < B;
         row += static_cast<long long>(gridDim.x) * warps) {
        float xv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
            xv[i] = i < per ? __bfloat162float(x[row * ldx + lane + 32 * i]) : 0.0f;
        float mine = -INFINITY;  // lane c keeps logit c
        for (int c = 0; c < C; ++c) {
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < per) acc += xv[i] * w_s[c * K + lane + 32 * i];
            acc = warp_sum(acc) + __ldg(b + c);
            if (lane == c) mine = acc;
        }
        const float mx = warp_max(mine);
        const float e = lane < C ? __expf(mine - mx) : 0.0f;
        const float den = warp_sum(e);
        if (lane < C) {
            logits[row * C + lane] = mine;
            if (probs) probs[row * C + lane] = e / den;
        }
    }
}

// ------------------------------------------------------------------ casts / layout helpers
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, long long ldx, int rows,
                                        int width4, __nv_bfloat16* __restrict__ y, long long ldy) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * width4) return;
    const int c = static_cast<int>(i % width4);
    const long long r = i / width4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + c);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(y + r * ldy + c * 4) = o;
}

__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                        int rows, int width8, float* __restrict__ y,
                                        long long ldy) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * width8) return;
    const int c = static_cast<int>(i % width8);
    const long long r = i / width8;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ldx) + c), f);
    float4* o = reinterpret_cast<float4*>(y + r * ldy + c * 8);
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// 32x32 smem transpose per (image, pixel tile, channel tile)
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C,
                                    float* __restrict__ y) {
    __shared__ float t[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int p = p0 + j, c = c0 + threadIdx.x;
        t[j][threadIdx.x] =
            (p < HW && c < C) ? __bfloat162float(x[(static_cast<long long>(n) * HW + p) * C + c])
                              : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, p = p0 + threadIdx.x;
        if (p < HW && c < C) y[(static_cast<long long>(n) * C + c) * HW + p] = t[threadIdx.x][j];
    }
}

__global__ void fill_f32_kernel(float* __restrict__ y, long long n, float v) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) y[i] = v;
}

// ------------------------------------------------------------------ weight packing (one-time)
__global__ void pack_conv_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int Cout, int Cin,
                                    int k, __nv_bfloat16* __restrict__ w_out,
                                    float* __restrict__ bias_out) {
    const long long total = static_cast<long long>(Cout) * k * k * Cin;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci = static_cast<int>(i % Cin);
    long long r = i / Cin;
    const int s = static_cast<int>(r % k);
    r /= k;
    const int rr = static_cast<int>(r % k);
    const int co = static_cast<int>(r / k);
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    w_out[i] = __float2bfloat16_rn(
        w[((static_cast<long long>(co) * Cin + ci) * k + rr) * k + s] * scale);
    if (ci == 0 && s == 0 && rr == 0) bias_out[co] = beta[co] - mean[co] * scale;
}

__global__ void pack_concat_k_kernel(const __nv_bfloat16* __restrict__ w0, int k0,
                                     const __nv_bfloat16* __restrict__ w1, int k1,
                                     const float* __restrict__ b0, const float* __restrict__ b1, int rows,
                                     __nv_bfloat16* __restrict__ w_out, float* __restrict__ b_out) {
    const int kt = k0 + k1;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * kt) return;
    const int c = static_cast<int>(i % kt);
    const int r = static_cast<int>(i / kt);
    w_out[i] = c < k0 ? w0[static_cast<long long>(r) * k0 + c] : w1[static_cast<long long>(r) * k1 + (c - k0)];
    if (c == 0) b_out[r] = b0[r] + b1[r];
}

__global__ void pack_stem_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps,
                                    __nv_bfloat16* __restrict__ w_out,
                                    float* __restrict__ bias_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 64*7*32
    if (i >= 64 * 7 * 32) return;
    const int slot = i & 31;
    const int r = (i >> 5) % 7;
    const int co = i / (7 * 32);
    const int s = slot >> 2, c = slot & 3;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    float v = 0.0f;
    if (s < 7 && c < 3) v = w[((co * 3 + c) * 7 + r) * 7 + s] * scale;
    w_out[i] = __float2bfloat16_rn(v);
    if (slot == 0 && r == 0) bias_out[co] = beta[co] - mean[co] * scale;
}

__global__ void pack_pos_type_kernel(const float* __restrict__ pos, const float* __restrict__ type0,
                                     int S, int Hd, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(S) * Hd) return;
    out[i] = pos[i] + type0[i % Hd];
}

__global__ void pack_linear_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                   long long n, int cols, int rows, float scale,
                                   __nv_bfloat16* __restrict__ w_out, float* __restrict__ b_out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) w_out[i] = __float2bfloat16_rn(w[i] * scale);
    if (b && b_out && i < rows) b_out[i] = b[i] * scale;
}

// W_out[i][j] = sum_k Wo[i][k] * Wv[k][j];  b_out[i] = sum_k Wo[i][k] * bv[k] + bo[i]
__global__ void pack_premul_kernel(const float* __restrict__ Wo, const float* __restrict__ bo,
                                   const float* __restrict__ Wv, const float* __restrict__ bv, int D,
                                   __nv_bfloat16* __restrict__ w_out, float* __restrict__ b_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= D) return;
    float acc = 0.0f;
    for (int k = 0; k < D; ++k) acc += Wo[static_cast<long long>(i) * D + k] * Wv[static_cast<long long>(k) * D + j];
    w_out[static_cast<long long>(i) * D + j] = __float2bfloat16_rn(acc);
    if (j == 0) {
        float bacc = bo[i];
        for (int k = 0; k < D; ++k) bacc += Wo[static_cast<long long>(i) * D + k] * bv[k];
        b_out[i] = bacc;
    }
}

}  // namespace

// ====================================================================== host wrappers
int repack_images(const void* x, bool x_is_bf16, int N, int H, int W, __nv_bfloat16* xpad,
                  cudaStream_t s) {
    if (N <= 0) return 0;
    if (W % 4 != 0 || N > 65535) {
        set_last_error("repack_images: W %% 4 != 0 or more than 65535 images per pass (W=%d N=%d)", W, N);
        return -1;
    }
    const int quads = (H + 6) * ((W + 8) / 4);
    const dim3 grid((quads + 255) / 256, N);
    if (x_is_bf16)
        repack_images_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), H, W,
                                                                  reinterpret_cast<uint4*>(xpad));
    else
        repack_images_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), H, W,
                                                          reinterpret_cast<uint4*>(xpad));
    return check_launch("repack_images");
}

int maxpool3x3s2(const __nv_bfloat16* x, int N, int H, int W, int C, __nv_bfloat16* y,
                 cudaStream_t s) {
    if (C % 8 != 0 || H % 2 != 0 || W % 2 != 0) {
        set_last_error("maxpool3x3s2: need C %% 8 == 0 and even H, W (C=%d H=%d W=%d)", C, H, W);
        return -1;
    }
    const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
    if (total <= 0) return 0;
    maxpool3x3s2_kernel<<<blocks_for(total, 256), 256, 0, s>>>(
        reinterpret_cast<const uint4*>(x), N, H, W, C / 8, reinterpret_cast<uint4*>(y));
    return check_launch("maxpool3x3s2");
}

int global_avgpool(const __nv_bfloat16* x, int N, int HW, int C, __nv_bfloat16* y_bf16,
                   float* y_f32, cudaStream_t s) {
    if (C % 8 != 0) {
        set_last_error("global_avgpool: C %% 8 != 0 (C=%d)", C);
        return -1;
    }
    const long long total = static_cast<long long>(N) * (C / 8);
    if (total <= 0) return 0;
    global_avgpool_kernel<<<blocks_for(total, 128), 128, 0, s>>>(
        reinterpret_cast<const uint4*>(x), N, HW, C / 8, reinterpret_cast<uint4*>(y_bf16), y_f32);
    return check_launch("global_avgpool");
}

int layernorm_residual(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* residual,
                       long long ldr, const float* gamma, const float* beta, float eps, int rows,
                       int width, __nv_bfloat16* y_bf16, long long ldy, float* y_f32,
                       long long ldy32, cudaStream_t s, const int* dyn_rows) {
    if (rows <= 0) return 0;
    if (ldx % 8 != 0 || (residual && ldr % 8 != 0) || (y_bf16 && ldy % 8 != 0) ||
        (y_f32 && ldy32 % 4 != 0)) {
        set_last_error("layernorm_residual: row strides must keep 16-byte alignment");
        return -1;
    }
    const unsigned grid = blocks_for(rows, 8);
#define MRD_LN(NC)                                                                              \
    layernorm_kernel<NC><<<grid, 256, 0, s>>>(x, ldx, residual, ldr, gamma, beta, eps, rows,    \
                                              y_bf16, ldy, y_f32, ldy32, dyn_rows)
    switch (width) {
        case 256: MRD_LN(1); break;
        case 512: MRD_LN(2); break;
        case 768: MRD_LN(3); break;
        case 1024: MRD_LN(4); break;
        default:
            set_last_error("layernorm_residual: unsupported width %d", width);
            return -1;
    }
#undef MRD_LN
    return check_launch("layernorm_residual");
}

int bert_embed_layernorm(const long long* ids, int B, int S, const __nv_bfloat16* word_emb,
                         const float* pos_type_emb, const float* gamma, const float* beta,
                         float eps, int vocab, __nv_bfloat16* y, cudaStream_t s, const int* row_tok,
                         const int* dyn_rows) {
    const long long tokens = static_cast<long long>(B) * S;
    if (tokens <= 0) return 0;
    bert_embed_ln_kernel<<<blocks_for(tokens, 8), 256, 0, s>>>(
        ids, static_cast<int>(tokens), S, word_emb, pos_type_emb, gamma, beta, eps, vocab, y, row_tok,
        dyn_rows);
    return check_launch("bert_embed_layernorm");
}

int compact_tokens(const void* mask, int mask_dtype, int B, int S, int keep_all, int* seq_off,
                   int* row_tok, float* row_bias, int* n_rows, int* scratch, cudaStream_t s) {
    if (B <= 0 || S <= 0) return 0;
    if (mask && (mask_dtype < MRD_DT_I64 || mask_dtype > MRD_DT_BF16)) {
        set_last_error("compact_tokens: unknown mask dtype code %d", mask_dtype);
        return -1;
    }
    compact_count_kernel<<<blocks_for(B, 8), 256, 0, s>>>(mask, mask_dtype, B, S, keep_all, scratch);
    compact_scan_kernel<<<1, 1024, 0, s>>>(scratch, B, seq_off, n_rows);
    compact_fill_kernel<<<blocks_for(B, 8), 256, 0, s>>>(mask, mask_dtype, B, S, keep_all, seq_off,
                                                        row_tok, row_bias);
    return check_launch("compact_tokens");
}

int gather_cls_rows(const __nv_bfloat16* x, const int* seq_off, int B, int width,
                    __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t s) {
    if (B <= 0) return 0;
    if (width % 8 != 0) {
        set_last_error("gather_cls_rows: width %% 8 != 0");
        return -1;
    }
    const long long total = static_cast<long long>(B) * (width / 8);
    gather_cls_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, seq_off, B, width / 8,
                                                           reinterpret_cast<uint4*>(y_bf16), y_f32);
    return check_launch("gather_cls_rows");
}

int mask_to_bias(const void* mask, int mask_dtype, int B, int S, float* bias, cudaStream_t s) {
    const long long n = static_cast<long long>(B) * S;
    if (n <= 0) return 0;
    if (mask == nullptr) return fill_f32(bias, n, 0.0f, s);
    if (mask_dtype < MRD_DT_I64 || mask_dtype > MRD_DT_BF16) {
        set_last_error("mask_to_bias: unknown mask dtype code %d", mask_dtype);
        return -1;
    }
    mask_to_bias_kernel<<<blocks_for(n, 256), 256, 0, s>>>(mask, mask_dtype, n, bias);
    return check_launch("mask_to_bias");
}

int head_logits_softmax(const __nv_bfloat16* x, long long ldx, const float* W, const float* b,
                        int B, int K, int C, float* logits, float* probs, cudaStream_t s) {
    if (B <= 0) return 0;
    if (C < 1 || C > 32 || K % 32 != 0 || K > 1024 || K <= 0) {
        set_last_error("head_logits_softmax: unsupported C=%d K=%d", C, K);
        return -1;
    }
    const size_t smem = static_cast<size_t>(C) * K * sizeof(float);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaFuncSetAttribute(head_logits_softmax_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
        attr_set = true;
    }
    unsigned grid = blocks_for(B, 8);
    if (grid > 592) grid = 592;
    head_logits_softmax_kernel<<<grid, 256, smem, s>>>(x, ldx, W, b, B, K, C, logits, probs);
    return check_launch("head_logits_softmax");
}

int cast_f32_to_bf16(const float* x, long long ldx, int rows, int width, __nv_bfloat16* y,
                     long long ldy, cudaStream_t s) {
    if (rows <= 0 || width <= 0) return 0;
    if (width % 4 != 0 || ldx % 4 != 0 || ldy % 4 != 0) {
        set_last_error("cast_f32_to_bf16: width/strides must be multiples of 4");
        return -1;
    }
    const long long total = static_cast<long long>(rows) * (width / 4);
    cast_f32_to_bf16_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, ldx, rows, width / 4, y, ldy);
    return check_launch("cast_f32_to_bf16");
}

int cast_bf16_to_f32(const __nv_bfloat16* x, long long ldx, int rows, int width, float* y,
                     long long ldy, cudaStream_t s) {
    if (rows <= 0 || width <= 0) return 0;
    if (width % 8 != 0 || ldx % 8 != 0 || ldy % 4 != 0) {
        set_last_error("cast_bf16_to_f32: width/strides must keep 16-byte alignment");
        return -1;
    }
    const long long total = static_cast<long long>(rows) * (width / 8);
    cast_bf16_to_f32_kernel<<<blocks_for(total, 256), 256, 0, s>>>(x, ldx, rows, width / 8, y, ldy);
    return check_launch("cast_bf16_to_f32");
}

int nhwc_bf16_to_nchw_f32(const __nv_bfloat16* x, int N, int HW, int C, float* y, cudaStream_t s) {
    if (N <= 0) return 0;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
    nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, s>>>(x, HW, C, y);
    return check_launch("nhwc_bf16_to_nchw_f32");
}

int fill_f32(float* y, long long n, float v, cudaStream_t s) {
    if (n <= 0) return 0;
    fill_f32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(y, n, v);
    return check_launch("fill_f32");
}

int pack_conv_bn(const float* w, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, int Cout, int Cin, int k, __nv_bfloat16* w_out,
                 float* bias_out, cudaStream_t s) {
    const long long total = static_cast<long long>(Cout) * Cin * k * k;
    pack_conv_bn_kernel<<<blocks_for(total, 256), 256, 0, s>>>(w, gamma, beta, mean, var, eps, Cout,
                                                              Cin, k, w_out, bias_out);
    return check_launch("pack_conv_bn");
}

int pack_concat_k(const __nv_bfloat16* w0, int k0, const __nv_bfloat16* w1, int k1, const float* b0,
                  const float* b1, int rows, __nv_bfloat16* w_out, float* b_out, cudaStream_t s) {
    const long long total = static_cast<long long>(rows) * (k0 + k1);
    pack_concat_k_kernel<<<blocks_for(total, 256), 256, 0, s>>>(w0, k0, w1, k1, b0, b1, rows, w_out, b_out);
    return check_launch("pack_concat_k");
}

int pack_stem_bn(const float* w, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, __nv_bfloat16* w_out, float* bias_out,
                 cudaStream_t s) {
    pack_stem_bn_kernel<<<blocks_for(64 * 7 * 32, 256), 256, 0, s>>>(w, gamma, beta, mean, var, eps,
                                                                    w_out, bias_out);
    return check_launch("pack_stem_bn");
}

int pack_pos_type(const float* pos, const float* type0, int S, int Hd, float* out, cudaStream_t s) {
    const long long total = static_cast<long long>(S) * Hd;
    pack_pos_type_kernel<<<blocks_for(total, 256), 256, 0, s>>>(pos, type0, S, Hd, out);
    return check_launch("pack_pos_type");
}

int pack_linear(const float* w, const float* b, int rows, int cols, float scale,
                __nv_bfloat16* w_out, float* b_out, cudaStream_t s) {
    const long long n = static_cast<long long>(rows) * cols;
    pack_linear_kernel<<<blocks_for(n, 256), 256, 0, s>>>(w, b, n, cols, rows, scale, w_out, b_out);
    return check_launch("pack_linear");
}

int pack_premul_linear(const float* Wo, const float* bo, const float* Wv, const float* bv, int D,
                       __nv_bfloat16* w_out, float* b_out, cudaStream_t s) {
    dim3 grid((D + 127) / 128, D);
    pack_premul_kernel<<<grid, 128, 0, s>>>(Wo, bo, Wv, bv, D, w_out, b_out);
    return check_launch("pack_premul_linear");
}

}  // namespace mrd
