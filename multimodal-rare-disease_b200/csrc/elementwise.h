// Memory-bound kernels (K4/K5 in DESIGN.md): input repack, max/avg pooling, residual+LayerNorm,
// BERT embedding-sum+LayerNorm, attention-mask conversion, classifier tail, and the one-time weight
// packing kernels.  All are vectorised (16-byte) and coalesced; roofline = HBM bandwidth.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mrd {

// x: [N,3,H,W] fp32 or bf16 (reference layout) -> xpad: [N][H+6][W+8][4] bf16, pixel (h,w) at
// (h+3,w+3), zero border, channel 3 zero.  Feeds the stem's overlapping-window TMA view.
int repack_images(const void* x, bool x_is_bf16, int N, int H, int W, __nv_bfloat16* xpad,
                  cudaStream_t s);

// 3x3 stride-2 pad-1 max pooling on NHWC bf16 (TV:models/resnet.py:200,271).  C % 8 == 0.
int maxpool3x3s2(const __nv_bfloat16* x, int N, int H, int W, int C, __nv_bfloat16* y,
                 cudaStream_t s);

// Global average pool over HW positions (TV:models/resnet.py:205,278).  x: [N,HW,C] -> [N,C].
int global_avgpool(const __nv_bfloat16* x, int N, int HW, int C, __nv_bfloat16* y_bf16,
                   float* y_f32, cudaStream_t s);

// y = LayerNorm(x + residual) * gamma + beta, one warp per row, fp32 statistics
// (HF:models/bert/modeling_bert.py:294-298,352-356; src/fusion_model.py:274-276).
// width in {256,512,768,1024}; residual may be null; either output may be null.
// dyn_rows (optional, device): the number of rows actually processed, <= rows (token-packed BERT).
int layernorm_residual(const __nv_bfloat16* x, long long ldx, const __nv_bfloat16* residual,
                       long long ldr, const float* gamma, const float* beta, float eps, int rows,
                       int width, __nv_bfloat16* y_bf16, long long ldy, float* y_f32,
                       long long ldy32, cudaStream_t s, const int* dyn_rows = nullptr);

// word[ids] + (position + token_type[0]) -> LayerNorm (HF:models/bert/modeling_bert.py:72-112).
// row_tok / dyn_rows (optional, device): packed row r embeds token row_tok[r] (= b*S + j) and only
// *dyn_rows rows exist; without them row r is token r.
int bert_embed_layernorm(const long long* ids, int B, int S, const __nv_bfloat16* word_emb,
                         const float* pos_type_emb, const float* gamma, const float* beta,
                         float eps, int vocab, __nv_bfloat16* y, cudaStream_t s,
                         const int* row_tok = nullptr, const int* dyn_rows = nullptr);

// Token packing ("unpadding") for BERT: keeps token (b,j) iff mask[b,j] != 0 or j == 0 (the CLS row
// is always needed, src/text_encoder.py:118) - or every token when keep_all.  Padded positions never
// influence attended ones (HF:masking_utils.py:1001-1088 masks them as keys) and only the CLS row is
// read downstream, so dropping them changes no output of TextEncoder.forward.
//   seq_off[B+1]: first packed row of each sequence; row_tok[r]: b*S + j of packed row r;
//   row_bias[r]: 0 / -inf additive key bias of packed row r; n_rows[0] = seq_off[B].
// scratch: B ints.  mask may be null (all ones).
int compact_tokens(const void* mask, int mask_dtype, int B, int S, int keep_all, int* seq_off,
                   int* row_tok, float* row_bias, int* n_rows, int* scratch, cudaStream_t s);

// CLS rows of a packed hidden state: row seq_off[b] of x -> y_bf16[b], y_f32[b] (either optional).
int gather_cls_rows(const __nv_bfloat16* x, const int* seq_off, int B, int width,
                    __nv_bfloat16* y_bf16, float* y_f32, cudaStream_t s);

// attention_mask [B,S] (MRD_DT_* code) -> additive key bias (0 or -inf), the key-padding semantics
// of HF:masking_utils.py:1001-1088.
int mask_to_bias(const void* mask, int mask_dtype, int B, int S, float* bias, cudaStream_t s);

// logits = x W^T + b ; probs = softmax(logits)   (src/multimodal_classifier.py:58,166-167).
// x: [B,K] bf16 (row stride ldx), W: [C,K] fp32, C <= 32, K % 32 == 0, K <= 1024.
int head_logits_softmax(const __nv_bfloat16* x, long long ldx, const float* W, const float* b,
                        int B, int K, int C, float* logits, float* probs, cudaStream_t s);

// fp32 [rows,width] (row stride ldx) -> bf16 [rows,width] (row stride ldy).  width % 4 == 0.
int cast_f32_to_bf16(const float* x, long long ldx, int rows, int width, __nv_bfloat16* y,
                     long long ldy, cudaStream_t s);
// bf16 [rows,width] (row stride ldx) -> fp32 [rows,width] (row stride ldy).  width % 8 == 0.
int cast_bf16_to_f32(const __nv_bfloat16* x, long long ldx, int rows, int width, float* y,
                     long long ldy, cudaStream_t s);
// NHWC bf16 [N,HW,C] -> NCHW f32 [N,C,HW] (explainability accessor, src/cnn_encoder.py:200-226).
int nhwc_bf16_to_nchw_f32(const __nv_bfloat16* x, int N, int HW, int C, float* y, cudaStream_t s);
// fill n floats with v (cross-attention weights are identically 1, src/fusion_model.py:164).
int fill_f32(float* y, long long n, float v, cudaStream_t s);

// ---- one-time weight packing -------------------------------------------------------------
// conv weight [Cout,Cin,k,k] fp32 + BatchNorm(gamma,beta,mean,var,eps) -> [Cout][k][k][Cin] bf16
// with the BN scale folded in, and bias[Cout] = beta - mean*scale.
int pack_conv_bn(const float* w, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, int Cout, int Cin, int k, __nv_bfloat16* w_out,
                 float* bias_out, cudaStream_t s);
// w_out[r] = [w0[r][0:k0] | w1[r][0:k1]] (bf16 rows concatenated along K), b_out = b0 + b1: the weights of two 1x1
// convolutions whose outputs are summed (conv3 + downsample of a bottleneck) as one K-concatenated GEMM.
int pack_concat_k(const __nv_bfloat16* w0, int k0, const __nv_bfloat16* w1, int k1, const float* b0,
                  const float* b1, int rows, __nv_bfloat16* w_out, float* b_out, cudaStream_t s);
// stem weight [64,3,7,7] + BN -> [64][7][32] bf16 (tap row, 8 pixels x 4 channels; unused slots 0)
int pack_stem_bn(const float* w, const float* gamma, const float* beta, const float* mean,
                 const float* var, float eps, __nv_bfloat16* w_out, float* bias_out,
                 cudaStream_t s);
// pos_type[s][h] = position[s][h] + token_type[0][h]
int pack_pos_type(const float* pos, const float* type0, int S, int Hd, float* out, cudaStream_t s);
// w_out[r][c] = bf16(w[r][c] * scale), b_out[r] = b[r] * scale   (rows x cols; b/b_out may be null)
int pack_linear(const float* w, const float* b, int rows, int cols, float scale,
                __nv_bfloat16* w_out, float* b_out, cudaStream_t s);
// W_out = Wo * Wv (bf16), b_out = Wo * bv + bo   (all [D,D] / [D]); the length-1 cross attention.
int pack_premul_linear(const float* Wo, const float* bo, const float* Wv, const float* bv, int D,
                       __nv_bfloat16* w_out, float* b_out, cudaStream_t s);

}  // namespace mrd
