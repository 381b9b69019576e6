// Counter-based dropout masks.  keep(seed, site, idx) is a pure function of a 64-bit seed, a 32-bit
// site id (which dropout layer) and the element index, so the backward pass recomputes the mask of the
// forward pass instead of storing it (and a test can materialise it: mrd_dropout_mask).
// The reference draws its masks from torch's Philox stream (nn.Dropout in train mode, e.g.
// HF:models/bert/modeling_bert.py:110,297,355; src/fusion_model.py:165); the streams cannot coincide,
// only the distribution does: P(keep) = 1 - p, kept values scaled by 1/(1-p).
#pragma once

#include <stdint.h>

namespace mrd {

struct DropCfg {
    unsigned long long seed;
    unsigned int site;
    unsigned int thresh;   // keep iff hash < thresh; 0 = dropout disabled (keep everything, scale 1)
    float scale;           // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t drop_hash(unsigned long long seed, unsigned int site,
                                                       unsigned long long idx) {
    // splitmix64 finaliser over (seed, site, idx)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull) +
                           0xD6E8FEB86659FD93ull * (static_cast<unsigned long long>(site) + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint32_t>(z >> 32);
}

__host__ __device__ __forceinline__ bool drop_keep(const DropCfg& d, unsigned long long idx) {
    return d.thresh == 0u || drop_hash(d.seed, d.site, idx) < d.thresh;
}

inline DropCfg make_drop(unsigned long long seed, unsigned int site, double p) {
    DropCfg d;
    d.seed = seed;
    d.site = site;
    if (p <= 0.0) {
        d.thresh = 0u;
        d.scale = 1.0f;
    } else {
        double t = (1.0 - p) * 4294967296.0;
        if (t < 1.0) t = 1.0;
        if (t > 4294967295.0) t = 4294967295.0;
        d.thresh = static_cast<unsigned int>(t);
        d.scale = static_cast<float>(1.0 / (1.0 - p));
    }
    return d;
}

}  // namespace mrd
