// conv_chain_kernel: conv3 (+downsample) + identity + ReLU of bottleneck i and conv1 + ReLU of bottleneck i+1 for
// the same 128-pixel tiles, as two tcgen05 pipelines inside one persistent CTA (see conv_chain.h).
//
//   warp 0        phase-0 TMA producer   A = conv2 output (and the block input when the downsample is fused), B = W1
//   warp 1        phase-0 MMA issuer     128 x 128 x 16 tcgen05.mma into TMEM columns [0, 256) (two accumulators);
//                                        owns the TMEM allocation (512 columns)
//   warps 2..9    epilogue of BOTH phases (TMEM lane quarter = warp % 4, two warps per quarter split 64 columns):
//                                        +bias (+identity) -> ReLU -> bf16 -> swizzled staging -> TMA store
//   warp 10       identity loader        TMA ring ahead of the phase-0 epilogue
//   warp 11       phase-1 TMA producer   A = the y tile this CTA stored `lag` tiles ago (an L2 hit), B = W2
//   warp 12       phase-1 MMA issuer     128 x BN2 x 16 into TMEM columns [256, 256 + 2*BN2)
//
// Hand-off of a y tile from phase 0 to phase 1: the epilogue thread that issues the TMA stores waits (one tile
// later, so the wait is free) until the bulk groups of tile j have COMPLETED, then arrives on y_ready[j & 3]; the
// phase-1 producer waits on it before its first load of tile j.  Both sides go through the async proxy and the
// data never leaves L2 in between.
// Reference ops replaced: TV:models/resnet.py:143-163 (Bottleneck.forward) across two consecutive blocks.

#include "conv_chain.h"

#include <stdio.h>
#include <string.h>

#include "gemm_conv.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace mrd {

namespace {

constexpr int kBN1 = 128;
constexpr int kThreads = 13 * 32;
constexpr int kEpi = 256;
constexpr int kSub = 128 * 128;      // one 128-row x 64-column bf16 sub-tile
constexpr int kASt = 128 * 128;      // A stage: 128 rows x 64 K
constexpr int kSmemMax = 232448;
constexpr int kBars = 1024;
constexpr int kAcc2Col = 256;        // first TMEM column of the phase-1 accumulators
// store staging buffers (measured: 4 instead of 2 changes nothing - the store stream is not what limits the
// memory-bound layers; the shared memory is better spent on the identity ring)
constexpr int kStg = 2;

int g_lag = 2, g_hints = 7;

struct Tile {
    int w0, h0, n0;
};
__device__ __forceinline__ Tile decode(const ChainParams& p, int m) {
    Tile t;
    const int tw_i = m % p.tiles_w;
    const int r = m / p.tiles_w;
    t.w0 = tw_i * p.tw;
    t.h0 = (r % p.tiles_h) * p.th;
    t.n0 = (r / p.tiles_h) * p.nb;
    return t;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// at most G bulk groups of this thread may still be pending (full completion, not just the smem reads)
__device__ __forceinline__ void store_wait_pending(int G) {
    switch (G) {
        case 1: tma_store_wait_all<1>(); break;
        case 2: tma_store_wait_all<2>(); break;
        case 4: tma_store_wait_all<4>(); break;
        case 8: tma_store_wait_all<8>(); break;
        case 16: tma_store_wait_all<16>(); break;
        case 32: tma_store_wait_all<32>(); break;
        default: tma_store_wait_all<0>(); break;
    }
}

template <int BN2>
__global__ void __launch_bounds__(kThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
    constexpr int kB1St = kBN1 * 128;
    constexpr int kB2St = BN2 * 128;
    constexpr int NSUB1 = kBN1 / 64, NSUB2 = BN2 / 64;
    const int S1 = p.stages1, S2 = p.stages2, RING = p.ring;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a1_s = base;
    const uint32_t b1_s = a1_s + S1 * kASt;
    const uint32_t a2_s = b1_s + S1 * kB1St;
    const uint32_t b2_s = a2_s + S2 * kASt;
    const uint32_t st_s = b2_s + S2 * kB2St;
    const uint32_t ring_s = st_s + kStg * kSub;
    const uint32_t bar = ring_s + RING * kSub;
    uint8_t* st_gen = gen + (st_s - base);
    uint8_t* ring_gen = gen + (ring_s - base);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar - base) + 480);

    auto full1 = [&](int s) { return bar + 8u * s; };
    auto empty1 = [&](int s) { return bar + 64u + 8u * s; };
    auto full2 = [&](int s) { return bar + 128u + 8u * s; };
    auto empty2 = [&](int s) { return bar + 192u + 8u * s; };
    auto rfull = [&](int s) { return bar + 256u + 8u * s; };
    auto rempty = [&](int s) { return bar + 320u + 8u * s; };
    auto tfull1 = [&](int a) { return bar + 384u + 8u * a; };
    auto tempty1 = [&](int a) { return bar + 400u + 8u * a; };
    auto tfull2 = [&](int a) { return bar + 416u + 8u * a; };
    auto tempty2 = [&](int a) { return bar + 432u + 8u * a; };
    auto yready = [&](int i) { return bar + 448u + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int ph = 0; ph < 2; ++ph) {
            tma_prefetch_desc(&p.ph[ph].a_map[0]);
            tma_prefetch_desc(&p.ph[ph].a_map[1]);
            tma_prefetch_desc(&p.ph[ph].b_map);
            tma_prefetch_desc(&p.ph[ph].c_map);
        }
        tma_prefetch_desc(&p.ph[0].r_map);
        for (int s = 0; s < 8; ++s) {
            mbar_init(full1(s), 1);
            mbar_init(empty1(s), 1);
            mbar_init(full2(s), 1);
            mbar_init(empty2(s), 1);
            mbar_init(rfull(s), 1);
            mbar_init(rempty(s), 8);   // one arrival per epilogue warp
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull1(a), 1);
            mbar_init(tempty1(a), 8);
            mbar_init(tfull2(a), 1);
            mbar_init(tempty2(a), 8);
        }
        for (int i = 0; i < 4; ++i) mbar_init(yready(i), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(bar + 480);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int bid = static_cast<int>(blockIdx.x), nblk = static_cast<int>(gridDim.x);
    const int n_my = p.m_tiles > bid ? (p.m_tiles - bid + nblk - 1) / nblk : 0;
    const int box_rows = p.tw * p.th * p.nb;
    const uint32_t a_bytes = static_cast<uint32_t>(box_rows) * 128u;
    const ChainPhase& P1 = p.ph[0];
    const ChainPhase& P2 = p.ph[1];

    if (warp == 0) {
        // ------------------------------------------------------------ phase-0 producer
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_my; ++j) {
            const Tile t = decode(p, bid + j * nblk);
            for (int n = 0; n < P1.n_tiles; ++n) {
                for (int ks = 0; ks < P1.num_k; ++ks) {
                    mbar_wait(empty1(stage), phase ^ 1u);
                    if (lane == 0) {
                        mbar_expect_tx(full1(stage), a_bytes + kB1St);
                        const int second = (P1.kc_split && ks >= P1.kc_split) ? 1 : 0;
                        const int kc = ks - (second ? P1.kc_split : 0);
                        if (p.hints & 1)
                            tma_load_4d_hint(&P1.a_map[second], full1(stage), a1_s + stage * kASt, kc * 64, t.w0, t.h0,
                                             t.n0, l2_policy_evict_first());
                        else
                            tma_load_4d(&P1.a_map[second], full1(stage), a1_s + stage * kASt, kc * 64, t.w0, t.h0, t.n0);
                        tma_load_2d(&P1.b_map, full1(stage), b1_s + stage * kB1St, ks * 64, n * kBN1);
                    }
                    __syncwarp();
                    if (++stage == S1) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ phase-0 MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, kBN1, 0, 0);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_my; ++j) {
            for (int n = 0; n < P1.n_tiles; ++n, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty1(acc), ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem + acc * kBN1;
                for (int ks = 0; ks < P1.num_k; ++ks) {
                    mbar_wait(full1(stage), phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t ad = make_smem_desc(a1_s + stage * kASt, 0, 1024, 2);
                        const uint64_t bd = make_smem_desc(b1_s + stage * kB1St, 0, 1024, 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
                        umma_commit(empty1(stage));
                        if (ks == P1.num_k - 1) umma_commit(tfull1(acc));
                    }
                    __syncwarp();
                    if (++stage == S1) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 10) {
        // ------------------------------------------------------------ identity loader (phase 0)
        if (P1.has_res) {
            int slot = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_my; ++j) {
                const Tile t = decode(p, bid + j * nblk);
                for (int n = 0; n < P1.n_tiles; ++n)
                    for (int sub = 0; sub < NSUB1; ++sub) {
                        mbar_wait(rempty(slot), phase ^ 1u);
                        if (lane == 0) {
                            mbar_expect_tx(rfull(slot), a_bytes);
                            if (p.hints & 1)
                                tma_load_4d_hint(&P1.r_map, rfull(slot), ring_s + slot * kSub, n * kBN1 + sub * 64, t.w0,
                                                 t.h0, t.n0, l2_policy_evict_first());
                            else
                                tma_load_4d(&P1.r_map, rfull(slot), ring_s + slot * kSub, n * kBN1 + sub * 64, t.w0,
                                            t.h0, t.n0);
                        }
                        __syncwarp();
                        if (++slot == RING) { slot = 0; phase ^= 1u; }
                    }
            }
        }
    } else if (warp == 11) {
        // ------------------------------------------------------------ phase-1 producer
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_my; ++j) {
            const Tile t = decode(p, bid + j * nblk);
            mbar_wait(yready(j & 3), static_cast<uint32_t>(j >> 2) & 1u);   // tile j of y is complete in memory
            fence_proxy_async_all();
            for (int n = 0; n < P2.n_tiles; ++n) {
                for (int ks = 0; ks < P2.num_k; ++ks) {
                    mbar_wait(empty2(stage), phase ^ 1u);
                    if (lane == 0) {
                        mbar_expect_tx(full2(stage), a_bytes + kB2St);
                        if (p.hints & 4)
                            tma_load_4d_hint(&P2.a_map[0], full2(stage), a2_s + stage * kASt, ks * 64, t.w0, t.h0, t.n0,
                                             l2_policy_evict_first());
                        else
                            tma_load_4d(&P2.a_map[0], full2(stage), a2_s + stage * kASt, ks * 64, t.w0, t.h0, t.n0);
                        tma_load_2d(&P2.b_map, full2(stage), b2_s + stage * kB2St, ks * 64, n * BN2);
                    }
                    __syncwarp();
                    if (++stage == S2) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 12) {
        // ------------------------------------------------------------ phase-1 MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, BN2, 0, 0);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_my; ++j) {
            for (int n = 0; n < P2.n_tiles; ++n, ++it) {
                const int acc = it & 1;
                mbar_wait(tempty2(acc), ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem + kAcc2Col + acc * BN2;
                for (int ks = 0; ks < P2.num_k; ++ks) {
                    mbar_wait(full2(stage), phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t ad = make_smem_desc(a2_s + stage * kASt, 0, 1024, 2);
                        const uint64_t bd = make_smem_desc(b2_s + stage * kB2St, 0, 1024, 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
                        umma_commit(empty2(stage));
                        if (ks == P2.num_k - 1) umma_commit(tfull2(acc));
                    }
                    __syncwarp();
                    if (++stage == S2) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue of both phases (warps 2..9)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int epi_tid = threadIdx.x - 64;
        int q = 0;                       // sub-tiles produced so far (staging buffer = q % kStg)
        int it1 = 0, it2 = 0;            // accumulator uses of the two phases
        int rslot = 0;
        uint32_t rphase = 0;

        // one (tile, n-tile) of one phase: nsub 64-column sub-tiles out of `acc_col`
        auto item = [&](const ChainPhase& P, const Tile& t, int n_idx, int block_n, int nsub, uint32_t acc_col,
                        uint32_t tfull, uint32_t tempty, uint32_t tphase, bool with_res, bool keep_in_l2) {
            for (int sub = 0; sub < nsub; ++sub, ++q) {
                const uint32_t buf = static_cast<uint32_t>(q) % kStg;
                const int col0 = n_idx * block_n + sub * 64 + half * 32;
                if (sub == 0) {
                    mbar_wait(tfull, tphase);
                    tc_fence_after();
                }
                uint32_t v[32];
                tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + sub * 64 + half * 32, v);
                tmem_ld_wait();
                if (sub == nsub - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty);
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.bias + col0 + j));
                    f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
                    f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                    f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                }
                if (with_res) {
                    mbar_wait(rfull(rslot), rphase);
                    const uint8_t* r_row = ring_gen + rslot * kSub + row * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = (half * 4 + c) ^ (row & 7);
                        const uint4 r4 = *reinterpret_cast<const uint4*>(r_row + chunk * 16);
                        const float2 r0 = unpack_bf16(r4.x), r1 = unpack_bf16(r4.y), r2 = unpack_bf16(r4.z),
                                     r3 = unpack_bf16(r4.w);
                        f[c * 8 + 0] += r0.x; f[c * 8 + 1] += r0.y;
                        f[c * 8 + 2] += r1.x; f[c * 8 + 3] += r1.y;
                        f[c * 8 + 4] += r2.x; f[c * 8 + 5] += r2.y;
                        f[c * 8 + 6] += r3.x; f[c * 8 + 7] += r3.y;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rempty(rslot));
                    if (++rslot == RING) { rslot = 0; rphase ^= 1u; }
                }
                if (P.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                }
                // staging buffer `buf` was the source of the TMA store kStg sub-tiles ago: it must have been read
                if (epi_tid == 0) tma_store_wait_read<kStg - 1>();
                named_bar_sync(2, kEpi);
                uint8_t* st_row = st_gen + buf * kSub + row * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 o;
                    o.x = pack_bf16(f[c * 8 + 0], f[c * 8 + 1]);
                    o.y = pack_bf16(f[c * 8 + 2], f[c * 8 + 3]);
                    o.z = pack_bf16(f[c * 8 + 4], f[c * 8 + 5]);
                    o.w = pack_bf16(f[c * 8 + 6], f[c * 8 + 7]);
                    const int chunk = (half * 4 + c) ^ (row & 7);
                    *reinterpret_cast<uint4*>(st_row + chunk * 16) = o;
                }
                fence_proxy_async_smem();
                named_bar_sync(1, kEpi);
                if (epi_tid == 0) {
                    if (keep_in_l2)   // the y tile is read back by phase 1 a few microseconds from now
                        tma_store_4d_hint(&P.c_map, st_s + buf * kSub, n_idx * block_n + sub * 64, t.w0, t.h0, t.n0,
                                          l2_policy_evict_last());
                    else
                        tma_store_4d(&P.c_map, st_s + buf * kSub, n_idx * block_n + sub * 64, t.w0, t.h0, t.n0);
                    tma_store_commit();
                }
            }
        };
        auto phase1_tile = [&](int j) {
            const Tile t = decode(p, bid + j * nblk);
            for (int n = 0; n < P2.n_tiles; ++n, ++it2) {
                const int acc = it2 & 1;
                item(P2, t, n, BN2, NSUB2, kAcc2Col + acc * BN2, tfull2(acc), tempty2(acc), (it2 >> 1) & 1u, false, false);
            }
        };
        const int groups1 = P1.n_tiles * NSUB1;   // bulk groups one tile of y consists of
        for (int j = 0; j < n_my; ++j) {
            const Tile t = decode(p, bid + j * nblk);
            for (int n = 0; n < P1.n_tiles; ++n, ++it1) {
                const int acc = it1 & 1;
                item(P1, t, n, kBN1, NSUB1, acc * kBN1, tfull1(acc), tempty1(acc), (it1 >> 1) & 1u, P1.has_res != 0,
                     (p.hints & 2) != 0);
            }
            // everything older than this tile's own groups has completed -> the previous y tile is in memory
            if (j >= 1 && epi_tid == 0) {
                store_wait_pending(groups1);
                fence_proxy_async_all();
                mbar_arrive(yready((j - 1) & 3));
            }
            if (j >= p.lag) phase1_tile(j - p.lag);
        }
        if (n_my > 0 && epi_tid == 0) {
            tma_store_wait_all<0>();
            fence_proxy_async_all();
            mbar_arrive(yready((n_my - 1) & 3));
        }
        for (int j = (n_my > p.lag ? n_my - p.lag : 0); j < n_my; ++j) phase1_tile(j);
        if (epi_tid == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

template <int BN2>
int launch_bn2(const ChainLaunch* g, cudaStream_t stream) {
    static bool attr_set = false;
    auto kfn = conv_chain_kernel<BN2>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
        if (e != cudaSuccess) {
            set_last_error("conv_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -static_cast<int>(e);
        }
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = g->smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, g->p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("conv_chain_kernel<%d> launch: %s", BN2, cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return 0;
}

// 4-D view of an NHWC activation seen through (stride s) sampling, boxes of tw x th x nb output pixels
int map_nhwc(CUtensorMap* m, const __nv_bfloat16* x, int C, int N, int Ho, int Wo, int s, int pad, int tw, int th,
             int nb) {
    const int W = Wo * s + 2 * pad, H = Ho * s + 2 * pad;
    const __nv_bfloat16* x0 = x + (static_cast<long long>(pad) * W + pad) * C;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)s * C * 2, (uint64_t)s * W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)nb};
    return encode_tensor_map(m, x0, 2, 4, dims, str, box, 128);
}
// the same tensor as a flat [M, C] matrix, boxes of 128 rows
int map_flat(CUtensorMap* m, const __nv_bfloat16* x, int C, long long M) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)M, 1, 1};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * M, (uint64_t)C * 2 * M};
    uint32_t box[4] = {64, 128, 1, 1};
    return encode_tensor_map(m, x, 2, 4, dims, str, box, 128);
}
int map_weights(CUtensorMap* m, const __nv_bfloat16* w, int N, int K, int bn) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    return encode_tensor_map(m, w, 2, 2, dims, str, box, 128);
}

}  // namespace

bool conv_chain_supported(int C0, int C1, int Cout, int C2) {
    return C0 > 0 && C0 % 64 == 0 && C1 % 64 == 0 && Cout % 128 == 0 && Cout / 64 <= 32 &&
           (C2 == 64 || C2 == 128 || C2 == 256);
}

int plan_conv_chain(ChainLaunch* g, const __nv_bfloat16* X0, int C0, const __nv_bfloat16* X1, int C1, int s1,
                    const __nv_bfloat16* identity, int N, int Ho, int Wo, const __nv_bfloat16* W1, int Cout,
                    const float* bias1, __nv_bfloat16* Y, const __nv_bfloat16* W2, int C2, const float* bias2,
                    __nv_bfloat16* Z, int out_pad) {
    memset(g, 0, sizeof(*g));
    if (!X1) C1 = 0;
    if (!conv_chain_supported(C0, C1, Cout, C2) || (X1 && s1 != 1 && s1 != 2) || (X1 && identity) || N <= 0 ||
        Ho <= 0 || Wo <= 0) {
        set_last_error("plan_conv_chain: unsupported shape C0=%d C1=%d Cout=%d C2=%d stride=%d", C0, C1, Cout, C2, s1);
        return -1;
    }
    ChainParams& p = g->p;
    const long long M = static_cast<long long>(N) * Ho * Wo;
    if (M > 0x7fffffffLL) {
        set_last_error("plan_conv_chain: M too large");
        return -1;
    }
    const int bn2 = C2 == 64 ? 64 : 128;
    g->block_n2 = bn2;
    const bool boxes = out_pad != 0 || (X1 && s1 == 2);   // tiles must be pixel boxes; otherwise flat 128-row tiles
    if (boxes) {
        p.tw = Wo < 128 ? Wo : 128;
        p.th = 128 / p.tw;
        if (p.th > Ho) p.th = Ho;
        if (p.th < 1) p.th = 1;
        p.nb = 1;
        if (p.th == Ho && p.tw == Wo) {
            p.nb = 128 / (p.tw * p.th);
            if (p.nb < 1) p.nb = 1;
            if (p.nb > N) p.nb = N;
        }
        p.tiles_w = (Wo + p.tw - 1) / p.tw;
        p.tiles_h = (Ho + p.th - 1) / p.th;
        p.tiles_img = (N + p.nb - 1) / p.nb;
    } else {
        p.tw = 128; p.th = 1; p.nb = 1;
        p.tiles_w = static_cast<int>((M + 127) / 128); p.tiles_h = 1; p.tiles_img = 1;
    }
    const long long mt = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_img;
    p.m_tiles = static_cast<int>(mt);
    auto act_map = [&](CUtensorMap* m, const __nv_bfloat16* x, int C, int s, int pad) -> int {
        if (boxes) return map_nhwc(m, x, C, N, Ho, Wo, s, pad, p.tw, p.th, p.nb);
        return map_flat(m, x, C, M);
    };
    ChainPhase& a = p.ph[0];
    ChainPhase& b = p.ph[1];
    int rc = act_map(&a.a_map[0], X0, C0, 1, 0);
    if (rc) return rc;
    a.a_map[1] = a.a_map[0];
    if (X1) {
        rc = act_map(&a.a_map[1], X1, C1, s1, 0);
        if (rc) return rc;
    }
    rc = map_weights(&a.b_map, W1, Cout, C0 + C1, kBN1);
    if (rc) return rc;
    rc = act_map(&a.c_map, Y, Cout, 1, 0);
    if (rc) return rc;
    a.r_map = a.c_map;
    if (identity) {
        rc = act_map(&a.r_map, identity, Cout, 1, 0);
        if (rc) return rc;
    }
    a.bias = bias1;
    a.num_k = (C0 + C1) / 64;
    a.kc_split = X1 ? C0 / 64 : 0;
    a.n_tiles = Cout / kBN1;
    a.act = ACT_RELU;
    a.has_res = identity ? 1 : 0;

    rc = act_map(&b.a_map[0], Y, Cout, 1, 0);
    if (rc) return rc;
    b.a_map[1] = b.a_map[0];
    rc = map_weights(&b.b_map, W2, C2, Cout, bn2);
    if (rc) return rc;
    rc = act_map(&b.c_map, Z, C2, 1, out_pad);
    if (rc) return rc;
    b.r_map = b.c_map;
    b.bias = bias2;
    b.num_k = Cout / 64;
    b.kc_split = 0;
    b.n_tiles = C2 / bn2;
    b.act = ACT_RELU;
    b.has_res = 0;

    // shared memory: [S1 x (A + B1)] [S2 x (A + B2)] [kStg staging] [ring] [barriers]
    const int st1 = kASt + kBN1 * 128, st2 = kASt + bn2 * 128;
    const int avail = kSmemMax - 1024 - kBars - kStg * kSub;
    p.ring = identity ? 2 : 0;
    p.stages1 = 2;
    p.stages2 = 2;
    int left = avail - p.ring * kSub - p.stages1 * st1 - p.stages2 * st2;
    if (left < 0) {
        set_last_error("plan_conv_chain: shared memory budget exceeded");
        return -1;
    }
    // leftovers go to the identity ring first: the identity is the largest DRAM stream of the launch and the only
    // one whose in-flight depth is set here (every other operand is small or an L2 hit)
    while (identity && left >= kSub && p.ring < 6) { ++p.ring; left -= kSub; }
    while (left >= st2 && p.stages2 < 6) { ++p.stages2; left -= st2; }
    while (identity && left >= kSub && p.ring < 8) { ++p.ring; left -= kSub; }
    g->smem = 1024 + p.stages1 * st1 + p.stages2 * st2 + kStg * kSub + p.ring * kSub + kBars;
    p.lag = g_lag;
    p.hints = g_hints;
    const int sms = gemm_num_sms();
    g->grid = p.m_tiles < sms ? p.m_tiles : sms;
    g->flops = 2.0 * M * (static_cast<double>(Cout) * (C0 + C1) + static_cast<double>(C2) * Cout);
    g->bytes = 2.0 * (1.0 * M * C0 + (X1 ? 1.0 * M * C1 : 0.0) + (identity ? 1.0 * M * Cout : 0.0) + 1.0 * M * Cout +
                      1.0 * M * C2 + 1.0 * Cout * (C0 + C1) + 1.0 * C2 * Cout);
    return 0;
}

void conv_chain_set_tuning(int lag, int hints) {
    g_lag = lag < 1 ? 1 : (lag > 3 ? 3 : lag);
    g_hints = hints;
}

int launch_conv_chain(const ChainLaunch* g, cudaStream_t stream) {
    if (g->block_n2 == 64) return launch_bn2<64>(g, stream);
    return launch_bn2<128>(g, stream);
}

}  // namespace mrd
