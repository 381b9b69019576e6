"""Seeded synthetic weights and inputs shared by bench.py, smoke(), the tests and oracle/make_golden.py.

Weights: the drop-in MultimodalClassifier built under torch.manual_seed(seed) (CPU RNG, so the same
torch build gives the same tensors on every machine).  `sensitise` then derives a second weight set
from it that makes parity meaningful: non-trivial BatchNorm statistics so
the BN folding is exercised, and larger fusion/head weights so logits differ across samples/classes.
"""

from __future__ import annotations

import hashlib
from typing import Dict, Tuple

import torch


def build_model(seed: int = 0):
    from .config import Config
    from .multimodal_classifier import MultimodalClassifier

    torch.manual_seed(seed)
    model = MultimodalClassifier(Config(), random_init=True)
    return model.eval()


def sensitise(sd: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if not v.is_floating_point():
            out[k] = v
            continue
        if "cnn_encoder.backbone" in k and (".bn" in k or "downsample.1" in k):
            if k.endswith("running_mean"):
                v = 0.1 * torch.randn(v.shape, generator=g)
            elif k.endswith("running_var"):
                v = 0.5 + torch.rand(v.shape, generator=g)
            elif k.endswith("weight"):
                # keep the residual branch (bn3) small so activations stay O(1..10) through 16 blocks
                scale = 0.25 if ".bn3." in k else 1.0
                v = scale * (0.5 + torch.rand(v.shape, generator=g))
            elif k.endswith("bias"):
                v = 0.1 * torch.randn(v.shape, generator=g)
        elif "LayerNorm" in k or "layer_norm" in k:
            if k.endswith("weight"):
                v = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
            else:
                v = 0.1 * torch.randn(v.shape, generator=g)
        elif k.startswith("fusion.") and k.endswith("weight"):
            v = v * 2.0
        elif k.startswith("classifier.") and k.endswith("weight"):
            v = v * 4.0
        elif k.startswith("cnn_encoder.projection") and k.endswith("weight"):
            v = v * 2.0
        elif k.endswith(".bias") and ("encoder.layer" in k or k.startswith("fusion.") or k.startswith("classifier.")):
            v = v + 0.05 * torch.randn(v.shape, generator=g)
        out[k] = v
    return out


def train_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Weights of the training-step fixtures (oracle/make_golden_train.py): plain random init with the
    sensitised LayerNorm parameters."""
    plain = build_model(seed).state_dict()
    full = sensitise(plain, 1)
    return {k: (full[k] if ("LayerNorm" in k or "layer_norm" in k) else v.clone()) for k, v in plain.items()}


def checksum(sd: Dict[str, torch.Tensor]) -> str:
    """Order-independent digest of a state_dict's float tensors (bit-exact)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k]
        if v.is_floating_point():
            h.update(k.encode())
            h.update(v.detach().float().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_inputs(B: int, S: int, seed: int, lengths=None, H: int = 224, W: int = 224,
                vocab_hi: int = 28000) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """images randn [B,3,H,W]; ids randint(1, vocab_hi) with [CLS]=101 first and 0 at padded
    positions; mask[b,j] = j < lengths[b] (all ones when lengths is None) - the shapes of
    src/train.py:608-611 plus the padded variants of SURVEY.md 8(d)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, H, W, generator=g)
    ids = torch.randint(1, vocab_hi, (B, S), generator=g)
    mask = torch.ones(B, S, dtype=torch.long)
    if lengths is not None:
        for b, L in enumerate(lengths):
            mask[b, L:] = 0
            ids[b, L:] = 0
    ids[:, 0] = 101
    return images, ids, mask


ROW_BLOCK = 64


def make_global_rows(lo: int, hi: int, seq: int = 128, img: int = 224, len_lo: int = 16, base_seed: int = 1234,
                     vocab: int = 28996) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Rows [lo, hi) of ONE global synthetic batch that does not depend on how it is sharded: block k (64
    consecutive samples) is drawn from its own generator seeded with (base_seed, k), so rank r of N holding
    samples [lo, hi) sees exactly the rows a single process would (bench.py: the gathered logits are
    bit-identical at N = 1, 2, 4, 8).  Same distribution as SURVEY.md 8(d) cfg 3/4: randn images, ids in
    [1, vocab) with [CLS] = 101 first and 0 on the padded tail, prefix masks with L ~ U{len_lo..seq}."""
    im, idl, ml = [], [], []
    for k in range(lo // ROW_BLOCK, -(-hi // ROW_BLOCK) if hi > lo else 0):
        g = torch.Generator().manual_seed(base_seed * 1000003 + k)
        images = torch.randn(ROW_BLOCK, 3, img, img, generator=g)
        ids = torch.randint(1, vocab, (ROW_BLOCK, seq), generator=g)
        lengths = torch.randint(len_lo, seq + 1, (ROW_BLOCK,), generator=g)
        mask = (torch.arange(seq).unsqueeze(0) < lengths.unsqueeze(1)).long()
        ids = ids * mask
        ids[:, 0] = 101
        a, b = max(lo, k * ROW_BLOCK) - k * ROW_BLOCK, min(hi, (k + 1) * ROW_BLOCK) - k * ROW_BLOCK
        im.append(images[a:b]); idl.append(ids[a:b]); ml.append(mask[a:b])
    if not im:
        return (torch.empty(0, 3, img, img), torch.empty(0, seq, dtype=torch.long),
                torch.empty(0, seq, dtype=torch.long))
    return torch.cat(im), torch.cat(idl), torch.cat(ml)


def tensor_digest(t: torch.Tensor) -> str:
    """sha256 of a tensor's fp32 bytes (bit-exact comparison of outputs across runs / world sizes)."""
    return hashlib.sha256(t.detach().float().cpu().contiguous().numpy().tobytes()).hexdigest()
