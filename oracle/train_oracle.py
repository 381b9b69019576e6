"""ORACLE (test infrastructure, NOT product code): the reference's TRAINING step restated on a plain
state_dict with torch.nn.functional ops, differentiable through torch.autograd.

Only tests/ may import this module.  What it restates (reference repo paths; TV: / HF: as in
forward_oracle.py):
  * train_forward   MultimodalClassifier.forward in train mode (src/multimodal_classifier.py:131-177):
                    dropout after the CNN projection's ReLU (src/cnn_encoder.py:46-51), BertEmbeddings /
                    BertSelfOutput / BertOutput dropouts (HF:110,297,355), dropout on the attention
                    probabilities (HF:168-207 -> sdpa dropout_p), TextEncoder.dropout on the CLS vector
                    (src/text_encoder.py:118-124), dropout on the length-1 cross-attention weights
                    (src/fusion_model.py:164-165), the fusion MLP dropout (src/fusion_model.py:232-237)
                    and the head dropouts (src/multimodal_classifier.py:44-56).  Dropout masks are
                    INPUTS (multiplicative tensors, already scaled by 1/(1-p)) because the reference's
                    Philox stream cannot be reproduced by another implementation; masks=None is p = 0.
                    BatchNorm: running statistics (backbone in eval mode) or batch statistics
                    (bn_train=True, F.batch_norm(training=True), TV:143-163 under model.train()).
  * train_step      src/train.py:247-321: CrossEntropyLoss -> backward -> clip_grad_norm_(1.0) ->
                    AdamW(lr, weight_decay) on every parameter that received a gradient.

Pinning: oracle/make_golden_train.py runs the UNMODIFIED reference model in train mode with every
dropout probability set to 0 (same seeded weights, inputs and labels) and stores its loss, gradients
and post-step parameters in tests/golden/train_*.pt; tests/test_oracle.py checks this file against
them.
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _bn(sd: SD, p: str, x: torch.Tensor, bn_train: bool, stats: Optional[dict], eps: float = 1e-5):
    if not bn_train:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                            sd[p + ".bias"], training=False, eps=eps)
    if stats is not None:  # what nn.BatchNorm2d would fold into its running buffers (momentum 0.1)
        n = x.numel() / x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var_unbiased = x.var(dim=(0, 2, 3), unbiased=False) * (n / max(n - 1.0, 1.0))
        stats[p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mean.detach()
        stats[p + ".running_var"] = 0.9 * sd[p + ".running_var"] + 0.1 * var_unbiased.detach()
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=eps)


def resnet50_pooled(sd: SD, x: torch.Tensor, bn_train: bool = False, stats: Optional[dict] = None,
                    p: str = "cnn_encoder.backbone.") -> torch.Tensor:
    """TV:266-282 up to avgpool + flatten -> [B,2048]."""
    x = F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(sd, p + "bn1", x, bn_train, stats))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for stage in range(1, 5):
        i = 0
        while f"{p}layer{stage}.{i}.conv1.weight" in sd:
            q = f"{p}layer{stage}.{i}."
            stride = 2 if (stage > 1 and i == 0) else 1
            identity = x
            y = F.relu(_bn(sd, q + "bn1", F.conv2d(x, sd[q + "conv1.weight"]), bn_train, stats))
            y = F.relu(_bn(sd, q + "bn2", F.conv2d(y, sd[q + "conv2.weight"], stride=stride, padding=1),
                           bn_train, stats))
            y = _bn(sd, q + "bn3", F.conv2d(y, sd[q + "conv3.weight"]), bn_train, stats)
            if q + "downsample.0.weight" in sd:
                identity = _bn(sd, q + "downsample.1", F.conv2d(x, sd[q + "downsample.0.weight"], stride=stride),
                               bn_train, stats)
            x = F.relu(y + identity)
            i += 1
    return x.mean(dim=(2, 3))


def train_forward(sd: SD, images: torch.Tensor, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                  masks: Optional[Dict[str, torch.Tensor]] = None, bn_train: bool = False,
                  stats: Optional[dict] = None, heads: int = 12, fusion_heads: int = 8,
                  bert_eps: float = 1e-12, fusion_eps: float = 1e-5, use_residual: bool = True) -> torch.Tensor:
    """-> logits [B,C].  masks keys: cnn_proj [B,512], emb / attn_out.{l} / ffn_out.{l} [B,S,768],
    attn.{l} [B,heads,S,S], text_out [B,768], i2t / t2i [B,fusion_heads], fusion_mlp [B,512], head.{j} [B,n]."""
    m = masks or {}

    def drop(x, key):
        return x * m[key] if key in m else x

    # ---- image branch (src/cnn_encoder.py:168-184)
    with torch.no_grad():   # frozen backbone: no gradient is needed below the projection
        feat = resnet50_pooled(sd, images, bn_train, stats)
    h = drop(F.relu(_lin(sd, "cnn_encoder.projection.0", feat)), "cnn_proj")
    img = _lin(sd, "cnn_encoder.projection.3", h)

    # ---- text branch (HF BertModel in train mode)
    p = "text_encoder.encoder."
    B, S = input_ids.shape
    e = p + "embeddings."
    x = (F.embedding(input_ids, sd[e + "word_embeddings.weight"], padding_idx=0)
         + sd[e + "token_type_embeddings.weight"][0]
         + sd[e + "position_embeddings.weight"][:S].unsqueeze(0))
    Hd = x.shape[-1]
    x = drop(F.layer_norm(x, (Hd,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], bert_eps), "emb")
    bias = None
    if attention_mask is not None:
        bias = torch.zeros(B, 1, 1, S, dtype=x.dtype)
        bias.masked_fill_((attention_mask == 0).view(B, 1, 1, S), float("-inf"))
    d = Hd // heads
    i = 0
    while f"{p}encoder.layer.{i}.attention.self.query.weight" in sd:
        L = f"{p}encoder.layer.{i}."
        q = _lin(sd, L + "attention.self.query", x).view(B, S, heads, d).transpose(1, 2)
        k = _lin(sd, L + "attention.self.key", x).view(B, S, heads, d).transpose(1, 2)
        v = _lin(sd, L + "attention.self.value", x).view(B, S, heads, d).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(d)
        if bias is not None:
            scores = scores + bias
        probs = drop(torch.softmax(scores, dim=-1), f"attn.{i}")
        ctx = (probs @ v).transpose(1, 2).reshape(B, S, Hd)
        a = drop(_lin(sd, L + "attention.output.dense", ctx), f"attn_out.{i}")
        x = F.layer_norm(a + x, (Hd,), sd[L + "attention.output.LayerNorm.weight"],
                         sd[L + "attention.output.LayerNorm.bias"], bert_eps)
        f = F.gelu(_lin(sd, L + "intermediate.dense", x))
        f = drop(_lin(sd, L + "output.dense", f), f"ffn_out.{i}")
        x = F.layer_norm(f + x, (Hd,), sd[L + "output.LayerNorm.weight"], sd[L + "output.LayerNorm.bias"], bert_eps)
        i += 1
    txt = drop(x[:, 0, :], "text_out")

    # ---- fusion (src/fusion_model.py:245-291; CrossModalAttention.forward :116-182 in full)
    fp = "fusion.fusion_layer."

    def cross(name, query, kv, key):
        qq = _lin(sd, fp + name + ".query_proj", query.unsqueeze(1))
        kk = _lin(sd, fp + name + ".key_proj", kv.unsqueeze(1))
        vv = _lin(sd, fp + name + ".value_proj", kv.unsqueeze(1))
        hidden = qq.shape[-1]
        dd = hidden // fusion_heads
        qq = qq.view(B, 1, fusion_heads, dd).transpose(1, 2)
        kk = kk.view(B, 1, fusion_heads, dd).transpose(1, 2)
        vv = vv.view(B, 1, fusion_heads, dd).transpose(1, 2)
        w = torch.softmax((qq @ kk.transpose(-2, -1)) * dd ** -0.5, dim=-1)   # [B,heads,1,1]
        if key in m:
            w = w * m[key].view(B, fusion_heads, 1, 1)
        out = (w @ vv).transpose(1, 2).reshape(B, 1, hidden)
        return _lin(sd, fp + name + ".output_proj", out).squeeze(1)

    ip = _lin(sd, fp + "image_proj", img)
    tp = _lin(sd, fp + "text_proj", txt)
    ia = cross("image_to_text_attention", ip, tp, "i2t")
    ta = cross("text_to_image_attention", tp, ip, "t2i")
    hdim = ip.shape[-1]
    io = F.layer_norm(ip + ia if use_residual else ia, (hdim,), sd[fp + "layer_norm_image.weight"],
                      sd[fp + "layer_norm_image.bias"], fusion_eps)
    to = F.layer_norm(tp + ta if use_residual else ta, (hdim,), sd[fp + "layer_norm_text.weight"],
                      sd[fp + "layer_norm_text.bias"], fusion_eps)
    fh = drop(F.relu(_lin(sd, fp + "fusion.0", torch.cat([io, to], -1))), "fusion_mlp")
    fused = _lin(sd, fp + "fusion.3", fh)

    # ---- head (src/multimodal_classifier.py:73-83)
    cp = "classifier.classifier."
    idx = sorted(int(k[len(cp):].split(".")[0]) for k in sd if k.startswith(cp) and k.endswith(".weight"))
    y = fused
    for j, li in enumerate(idx):
        y = _lin(sd, f"{cp}{li}", y)
        if j + 1 < len(idx):
            y = drop(F.relu(y), f"head.{j}")
    return y


def trainable_names(sd: SD):
    """Parameters the default configuration trains (src/config.py:64 freezes the backbone); buffers and the
    unused BERT pooler excluded."""
    out = []
    for k, v in sd.items():
        if not v.is_floating_point() or k.startswith("cnn_encoder.backbone.") or ".pooler." in k:
            continue
        if k.endswith(("running_mean", "running_var")):
            continue
        out.append(k)
    return out


def loss_and_grads(sd: SD, images, input_ids, attention_mask, labels, masks=None, bn_train=False, stats=None):
    """CrossEntropyLoss(logits, labels).backward() (src/train.py:258-264,316) -> (loss, logits, {name: grad})."""
    names = trainable_names(sd)
    work = {k: (v.detach().clone().float().requires_grad_(k in set(names)) if v.is_floating_point() else v)
            for k, v in sd.items()}
    logits = train_forward(work, images.float(), input_ids, attention_mask, masks, bn_train, stats)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    grads = {k: (work[k].grad if work[k].grad is not None else torch.zeros_like(work[k])) for k in names}
    return loss.detach(), logits.detach(), grads


def clip_and_adamw(sd: SD, grads: Dict[str, torch.Tensor], lr: float = 5e-5, weight_decay: float = 0.05,
                   max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8) -> Dict[str, torch.Tensor]:
    """clip_grad_norm_(parameters, 1.0) (src/train.py:317-320) then the FIRST AdamW step
    (src/train.py:193-198; torch defaults for betas/eps): returns the updated parameters."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    out = {}
    for k, g in grads.items():
        g = g * coef
        p = sd[k].float() * (1.0 - lr * weight_decay)
        m = (1 - betas[0]) * g
        v = (1 - betas[1]) * g * g
        m_hat = m / (1 - betas[0])
        v_hat = v / (1 - betas[1])
        out[k] = p - lr * m_hat / (v_hat.sqrt() + eps)
    return out
