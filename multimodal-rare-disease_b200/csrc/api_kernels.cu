// C-ABI entry points for the individual kernels (unit tests, ncu captures, other hosts).
// All pointers are device pointers; `stream` is a cudaStream_t passed as void*.

#include <mrd_b200.h>

#include "elementwise.h"
#include "attention.h"
#include "conv_chain.h"
#include "gemm_conv.h"
#include "tma_host.h"

using namespace mrd;

extern "C" {

const char* mrd_last_error(void) { return get_last_error(); }

int mrd_gemm_bf16(const void* A, long long lda, int M, int K, const void* W, int N,
                  const float* bias, void* C, long long ldc, const void* residual,
                  long long ld_res, float* out_f32, long long ld_f32, int act, void* stream) {
    GemmLaunch g;
    int rc = plan_gemm(&g, static_cast<const __nv_bfloat16*>(A), lda, M, K,
                       static_cast<const __nv_bfloat16*>(W), N, bias,
                       static_cast<__nv_bfloat16*>(C), ldc,
                       static_cast<const __nv_bfloat16*>(residual), ld_res, out_f32, ld_f32, act);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_gemm_ln_bf16(const void* A, long long lda, int M, int K, const void* W, int N, const float* bias, void* C,
                     long long ldc, const void* residual, long long ld_res, const float* gamma, const float* beta,
                     float eps, void* stats_ws, void* stream) {
    GemmLaunch g;
    int rc = plan_gemm_ln(&g, static_cast<const __nv_bfloat16*>(A), lda, M, K, static_cast<const __nv_bfloat16*>(W), N,
                          bias, static_cast<__nv_bfloat16*>(C), ldc, static_cast<const __nv_bfloat16*>(residual), ld_res,
                          gamma, beta, eps, stats_ws);
    if (rc > 0) {
        set_last_error("mrd_gemm_ln_bf16: N=%d outside the fused LayerNorm kernel's range (512 / 768 / 1024)", N);
        return -1;
    }
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

long long mrd_gemm_ln_ws_bytes(int M) { return static_cast<long long>(gemm_ln_ws_bytes(M)); }

int mrd_gemm_splitk_f32(const void* A, long long lda, int M, int K, const void* W, int N, float* out_f32,
                        long long ld_f32, const int* dyn_k, void* stream) {
    GemmLaunch g;
    int rc = plan_gemm_splitk(&g, static_cast<const __nv_bfloat16*>(A), lda, M, K,
                              static_cast<const __nv_bfloat16*>(W), N, out_f32, ld_f32, dyn_k);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_conv2d_nhwc_bf16(const void* X, int N, int H, int W, int Cin, const void* Wt, int Cout,
                         int ksize, int stride, const float* bias, void* Y, const void* residual,
                         int act, int out_pad, void* stream) {
    GemmLaunch g;
    int rc = plan_conv(&g, static_cast<const __nv_bfloat16*>(X), N, H, W, Cin,
                       static_cast<const __nv_bfloat16*>(Wt), Cout, ksize, stride, bias,
                       static_cast<__nv_bfloat16*>(Y), static_cast<const __nv_bfloat16*>(residual),
                       act, out_pad);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_conv1x1_dual_bf16(const void* X0, int C0, const void* X1, int C1, int stride, int N, int Ho, int Wo,
                          const void* Wcat, int Cout, const float* bias, void* Y, int act, void* stream) {
    GemmLaunch g;
    int rc = plan_conv1x1_dual(&g, static_cast<const __nv_bfloat16*>(X0), C0, static_cast<const __nv_bfloat16*>(X1),
                               C1, stride, N, Ho, Wo, static_cast<const __nv_bfloat16*>(Wcat), Cout, bias,
                               static_cast<__nv_bfloat16*>(Y), act);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_conv_chain_bf16(const void* X0, int C0, const void* X1, int C1, int stride, const void* identity, int N,
                        int Ho, int Wo, const void* W1, int Cout, const float* bias1, void* Y, const void* W2,
                        int C2, const float* bias2, void* Z, int out_pad, void* stream) {
    typedef const __nv_bfloat16* cb;
    ChainLaunch g;
    int rc = plan_conv_chain(&g, static_cast<cb>(X0), C0, static_cast<cb>(X1), C1, stride, static_cast<cb>(identity), N,
                             Ho, Wo, static_cast<cb>(W1), Cout, bias1, static_cast<__nv_bfloat16*>(Y),
                             static_cast<cb>(W2), C2, bias2, static_cast<__nv_bfloat16*>(Z), out_pad);
    if (rc) return rc;
    return launch_conv_chain(&g, static_cast<cudaStream_t>(stream));
}

int mrd_conv3x3_flat_bf16(const void* Xpad, int N, int H, int W, int Cin, const void* Wt, int Cout,
                          const float* bias, void* Y, int act, void* stream) {
    GemmLaunch g;
    int rc = plan_conv3x3_flat(&g, static_cast<const __nv_bfloat16*>(Xpad), N, H, W, Cin,
                               static_cast<const __nv_bfloat16*>(Wt), Cout, bias,
                               static_cast<__nv_bfloat16*>(Y), act);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_stem_conv_bf16(const void* Xpad, int N, int H, int W, const void* Wst, const float* bias,
                       void* Y, int act, void* stream) {
    GemmLaunch g;
    int rc = plan_stem(&g, static_cast<const __nv_bfloat16*>(Xpad), N, H, W,
                       static_cast<const __nv_bfloat16*>(Wst), bias,
                       static_cast<__nv_bfloat16*>(Y), act);
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_stem_pool_bf16(const void* Xpad, int N, int H, int W, const void* Wst, const float* bias, void* P,
                       void* stream) {
    GemmLaunch g;
    int rc = plan_stem_pool(&g, static_cast<const __nv_bfloat16*>(Xpad), N, H, W,
                            static_cast<const __nv_bfloat16*>(Wst), bias, static_cast<__nv_bfloat16*>(P));
    if (rc) return rc;
    return launch_gemm(&g, static_cast<cudaStream_t>(stream));
}

int mrd_repack_images(const void* x_nchw, int img_dtype, int N, int H, int W, void* xpad,
                      void* stream) {
    if (img_dtype != MRD_DT_F32 && img_dtype != MRD_DT_BF16) {
        set_last_error("mrd_repack_images: images must be f32 or bf16");
        return -1;
    }
    return repack_images(x_nchw, img_dtype == MRD_DT_BF16, N, H, W,
                         static_cast<__nv_bfloat16*>(xpad), static_cast<cudaStream_t>(stream));
}

int mrd_maxpool3x3s2(const void* x, int N, int H, int W, int C, void* y, void* stream) {
    return maxpool3x3s2(static_cast<const __nv_bfloat16*>(x), N, H, W, C,
                        static_cast<__nv_bfloat16*>(y), static_cast<cudaStream_t>(stream));
}

int mrd_global_avgpool(const void* x, int N, int HW, int C, void* y_bf16, float* y_f32,
                       void* stream) {
    return global_avgpool(static_cast<const __nv_bfloat16*>(x), N, HW, C,
                          static_cast<__nv_bfloat16*>(y_bf16), y_f32,
                          static_cast<cudaStream_t>(stream));
}

int mrd_layernorm_residual(const void* x, long long ldx, const void* residual, long long ldr,
                           const float* gamma, const float* beta, float eps, int rows, int width,
                           void* y_bf16, long long ldy, float* y_f32, long long ldy32,
                           void* stream) {
    return layernorm_residual(static_cast<const __nv_bfloat16*>(x), ldx,
                              static_cast<const __nv_bfloat16*>(residual), ldr, gamma, beta, eps,
                              rows, width, static_cast<__nv_bfloat16*>(y_bf16), ldy, y_f32, ldy32,
                              static_cast<cudaStream_t>(stream));
}

int mrd_bert_embed_layernorm(const long long* ids, int B, int S, const void* word_emb,
                             const float* pos_type_emb, const float* gamma, const float* beta,
                             float eps, int vocab, void* y_bf16, void* stream) {
    return bert_embed_layernorm(ids, B, S, static_cast<const __nv_bfloat16*>(word_emb),
                                pos_type_emb, gamma, beta, eps, vocab,
                                static_cast<__nv_bfloat16*>(y_bf16),
                                static_cast<cudaStream_t>(stream));
}

int mrd_attention_bf16(const void* qkv, const float* mask_bias, int B, int S, int heads,
                       void* out, void* stream) {
    return attention_forward(static_cast<const __nv_bfloat16*>(qkv), mask_bias, nullptr, B, S, heads,
                             static_cast<__nv_bfloat16*>(out), static_cast<cudaStream_t>(stream));
}

int mrd_attention_varlen_bf16(const void* qkv, const float* row_bias, const int* seq_off, int B,
                              int max_len, int heads, long long total_rows, void* out, void* stream) {
    if (!seq_off) {
        set_last_error("mrd_attention_varlen_bf16: seq_off is required");
        return -1;
    }
    if (total_rows <= 0) {
        set_last_error("mrd_attention_varlen_bf16: total_rows must be the row count of qkv / out");
        return -1;
    }
    return attention_forward(static_cast<const __nv_bfloat16*>(qkv), row_bias, seq_off, B, max_len,
                             heads, static_cast<__nv_bfloat16*>(out),
                             static_cast<cudaStream_t>(stream), total_rows);
}

int mrd_attention_use_tcgen05(int on) {
    attention_set_tc(on != 0);
    return 0;
}

int mrd_compact_tokens(const void* mask, int mask_dtype, int B, int S, int keep_all, int* seq_off,
                       int* row_tok, float* row_bias, int* n_rows, int* scratch, void* stream) {
    return compact_tokens(mask, mask_dtype, B, S, keep_all, seq_off, row_tok, row_bias, n_rows,
                          scratch, static_cast<cudaStream_t>(stream));
}

int mrd_mask_to_bias(const void* mask, int mask_dtype, int B, int S, float* bias, void* stream) {
    return mask_to_bias(mask, mask_dtype, B, S, bias, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
