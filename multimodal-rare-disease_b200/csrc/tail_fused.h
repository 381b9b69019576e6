// K6: the whole batch-level tail of MultimodalClassifier.forward in ONE launch - AttentionFusion
// (src/fusion_model.py:245-291: image/text projections, the two length-1 cross attentions with their residuals, the
// two LayerNorms, the concat MLP) followed by ClassificationHead + softmax (src/multimodal_classifier.py:73-83,
// 166-167).  A CTA owns 64 samples; the activations never leave the SM: every Linear is a tcgen05 product whose A
// operand is the previous epilogue's bf16 output in shared memory (K-major, SWIZZLE_128B) and whose weights stream
// through a TMA ring, accumulators live in TMEM, LayerNorm / ReLU / the final GEMV + softmax are epilogues.
//
// Algebra used (weights are prepared once per load by pack_tail_big):
//   image_proj, text_proj and the cross attentions are all affine and a softmax over ONE key is 1
//   (src/fusion_model.py:138-164), so
//     pre_i = image_proj(img) + O_i2t(V_i2t(text_proj(txt)))  = [img | txt] . [ Wip | P_i2t Wtp ]^T + const
//     pre_t = text_proj(txt)  + O_t2i(V_t2i(image_proj(img))) = [img | txt] . [ P_t2i Wip | Wtp ]^T + const
//   with P = Wo Wv.  One K-concatenated product per LayerNorm input replaces four chained ones.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mrd {

constexpr int kTailMaxStages = 8;
enum { TAIL_EPI_ACT = 0, TAIL_EPI_LN = 1, TAIL_EPI_FINAL = 2 };

struct TailStage {
    CUtensorMap w_map;      // weights [rows][K] bf16 (nn.Linear layout), box {64, bn}
    const float* bias;      // [n] fp32
    const float* ln_g;      // TAIL_EPI_LN: LayerNorm weight / bias [n]
    const float* ln_b;
    float* out_f32;         // optional fp32 copy of the stage's result rows (row stride ld_f32), or nullptr
    int ld_f32;
    int k_chunks;           // K / 64
    int n;                  // output columns (<= 512, multiple of 64)
    int bn;                 // accumulator tile width (n % bn == 0, bn <= 256)
    int w_row0;             // first weight row of this stage inside w_map
    int a_chunk0;           // A operand: first 64-column chunk of the resident activation area, or -1 = streamed
                            // from the image / text embeddings
    int out_chunk0;         // where the bf16 result goes (activation-area chunk)
    int epi;                // TAIL_EPI_*
    int act;                // ACT_NONE / ACT_RELU / ACT_GELU (gemm_conv.h codes)
};

struct alignas(64) TailParams {
    CUtensorMap img_map, txt_map;   // [B][img_in] / [B][txt_in] bf16, box {64 columns, 64 rows}
    TailStage st[kTailMaxStages];
    int num_stages;
    int img_chunks;                 // streamed K chunks [0, img_chunks) come from img_map, the rest from txt_map
    int B;
    float ln_eps;
    const float* out_w;             // final Linear [C][h_last] fp32
    const float* out_b;             // [C]
    int h_last, C;
    float* logits;                  // [B][C] fp32 or nullptr
    float* probs;                   // [B][C] fp32 or nullptr
};

struct TailLaunch {
    TailParams p;
    int grid;
    double flops, bytes;
};

struct TailWeights {
    const __nv_bfloat16* w_big;     // [2F][img_in + txt_in] (pack_tail_big)
    const float* b_big;             // [2F]
    const float *ln_i_g, *ln_i_b, *ln_t_g, *ln_t_b;
    const __nv_bfloat16 *w1, *w2;   // fusion.0 [F][2F], fusion.3 [F][F]
    const float *b1, *b2;
    int num_hidden;                 // classifier hidden Linear layers (1..3)
    const __nv_bfloat16* wh[3];
    const float* bh[3];
    int hdim[3];
    int head_act;
    const float* out_w;             // [C][hdim[last]] fp32
    const float* out_b;
    int F, img_in, txt_in, C;
    float ln_eps;
};

bool tail_supported(int F, int img_in, int txt_in, int num_hidden, const int* hdim, int C);
// img / txt: [B][img_in] / [B][txt_in] bf16 embeddings.  logits / probs / fused_f32 are patched per call
// (launch_tail), everything else is fixed by the plan.
int plan_tail(TailLaunch* out, const TailWeights& w, const __nv_bfloat16* img, const __nv_bfloat16* txt, int B);
int launch_tail(const TailLaunch* g, float* logits, float* probs, float* fused_f32, int ld_fused, cudaStream_t stream);

// Weight preparation from the raw fp32 parameters (device pointers): w_big [2F][Ii + Ti] bf16, b_big [2F] fp32.
// scratch: 2*F*F + 2*F floats.  residual = FusionConfig.use_residual (src/fusion_model.py:268-276).
struct CrossRaw {
    const float *wv, *bv, *wo, *bo;   // value_proj / output_proj of one CrossModalAttention, [F][F] / [F]
};
int pack_tail_big(const float* wip, const float* bip, const float* wtp, const float* btp, const CrossRaw& i2t,
                  const CrossRaw& t2i, int F, int Ii, int Ti, int residual, float* scratch, __nv_bfloat16* w_big,
                  float* b_big, cudaStream_t stream);

}  // namespace mrd
