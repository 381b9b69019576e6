// fp32 check mode: the whole MultimodalClassifier forward in plain fp32 CUDA (SIMT FFMA, no bf16
// storage anywhere, BatchNorm un-folded), driven from the caller's raw fp32 parameters.  It is the
// "fp32 check mode" of BASELINE.json's north_star (logits within 1e-4 of the reference's own fp32
// forward): slow by design - a verification instrument for the fast bf16/tcgen05 path, selected with
// mrd_ctx_set_option(ctx, "fp32_check", 1).
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <unordered_map>

namespace mrd {

struct RawTensor {
    const float* p = nullptr;
    long long d[4] = {0, 0, 0, 0};
    long long numel() const {
        long long n = 1;
        for (int i = 0; i < 4; ++i)
            if (d[i] > 0) n *= d[i];
        return n;
    }
};
typedef std::unordered_map<std::string, RawTensor> RawTable;

struct Fp32Opts {
    int bert_heads = 12;
    float bert_ln_eps = 1e-12f, bn_eps = 1e-5f, fusion_ln_eps = 1e-5f;
    int fusion_heads = 8, fusion_residual = 1, head_act = 1;
};

struct Fp32Arena {
    void* base = nullptr;
    size_t bytes = 0, used = 0;
};

// Each stage mirrors one module forward of the reference; outputs are fp32 device buffers owned by
// the caller.  Stages return 0 or a negative code (message via mrd_last_error()).
int fp32_cnn_encoder(const RawTable& t, const Fp32Opts& o, Fp32Arena* ws, const float* images, int B,
                     int H, int W, float* emb, float* pooled, float* fmap, cudaStream_t s);
int fp32_text_encoder(const RawTable& t, const Fp32Opts& o, Fp32Arena* ws, const long long* ids,
                      const void* mask, int mask_dtype, int B, int S, float* cls, float* last_hidden,
                      float* all_hidden, cudaStream_t s);
int fp32_fusion(const RawTable& t, const Fp32Opts& o, Fp32Arena* ws, const float* img, const float* txt,
                int B, float* fused, float* attn_i2t, float* attn_t2i, cudaStream_t s);
int fp32_head(const RawTable& t, const Fp32Opts& o, Fp32Arena* ws, const float* x, int B, float* logits,
              float* probs, cudaStream_t s);
void fp32_arena_free(Fp32Arena* ws);

}  // namespace mrd
