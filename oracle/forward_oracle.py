"""ORACLE (test infrastructure, NOT product code): a CPU fp32 restatement of the reference's
MultimodalClassifier forward, written against a plain state_dict with torch.nn.functional ops only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; nothing under multimodal-rare-disease_b200/ does.

What it restates (reference repo paths; TV: = torchvision 0.26.0 models/resnet.py, HF: = transformers
5.5.0 models/bert/modeling_bert.py - the two un-vendored third-party packages the reference calls
into, pinned only by lower bounds in requirements.txt:5-11):
  * resnet50_backbone   TV:266-282 (_forward_impl), TV:143-163 (Bottleneck.forward, stride on the 3x3
                        conv = v1.5, TV:109-113), eval-mode BatchNorm (running statistics, eps 1e-5)
  * cnn_encoder         src/cnn_encoder.py:168-184, projection src/cnn_encoder.py:46-51 (eval: no dropout)
  * bert_encoder        HF:72-112 (embeddings: word + token_type[0] + position, LayerNorm eps 1e-12),
                        HF:168-207 + integrations/sdpa_attention.py:92 (softmax(QK^T/sqrt(64) + key
                        padding mask) V; mask semantics HF:masking_utils.py:1001-1088: key j is visible
                        iff attention_mask[b,j] != 0), HF:294-298, 339-342, 352-356, exact-erf GELU
                        (HF:activations.py:70-90)
  * text_encoder        src/text_encoder.py:95-127 (CLS row of the last hidden state, eval dropout)
  * attention_fusion    src/fusion_model.py:245-291 with CrossModalAttention.forward :116-182 restated
                        IN FULL (query/key projections, scaled scores, softmax over the single key),
                        not the algebraic shortcut the CUDA path uses
  * classification_head src/multimodal_classifier.py:73-83, softmax :167

Pinning: the reference has no golden vectors or value-asserting tests for this path (SURVEY.md 8(c):
"parity unpinned" by the reference itself), so this oracle is pinned against OUTPUTS OF THE REFERENCE
RUN IN THE BUILD CONTAINER: oracle/make_golden.py imports the unmodified reference modules from
/root/reference, loads the same seeded state_dict, and writes tests/golden/*.pt;
tests/test_oracle.py checks this file against those fixtures.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _bn(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=eps)


def _linear(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


# ---------------------------------------------------------------------------------- ResNet50
def resnet50_feature_map(sd: SD, x: torch.Tensor, p: str = "cnn_encoder.backbone.") -> torch.Tensor:
    """[B,3,H,W] -> layer4 output [B,2048,H/32,W/32]."""
    x = F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(sd, p + "bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for stage in range(1, 5):
        i = 0
        while f"{p}layer{stage}.{i}.conv1.weight" in sd:
            q = f"{p}layer{stage}.{i}."
            stride = 2 if (stage > 1 and i == 0) else 1
            identity = x
            y = F.relu(_bn(sd, q + "bn1", F.conv2d(x, sd[q + "conv1.weight"])))
            y = F.relu(_bn(sd, q + "bn2", F.conv2d(y, sd[q + "conv2.weight"], stride=stride, padding=1)))
            y = _bn(sd, q + "bn3", F.conv2d(y, sd[q + "conv3.weight"]))
            if q + "downsample.0.weight" in sd:
                identity = _bn(sd, q + "downsample.1",
                               F.conv2d(x, sd[q + "downsample.0.weight"], stride=stride))
            x = F.relu(y + identity)
            i += 1
    return x


def cnn_encoder(sd: SD, images: torch.Tensor) -> torch.Tensor:
    """src/cnn_encoder.py:168-184 -> [B,512]."""
    feat = resnet50_feature_map(sd, images).mean(dim=(2, 3))  # AdaptiveAvgPool2d(1) + flatten
    h = F.relu(_linear(sd, "cnn_encoder.projection.0", feat))
    return _linear(sd, "cnn_encoder.projection.3", h)


# ---------------------------------------------------------------------------------- BERT
def bert_encoder(sd: SD, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                 heads: int = 12, eps: float = 1e-12, p: str = "text_encoder.encoder.") -> torch.Tensor:
    """[B,S] ids -> last hidden state [B,S,768]."""
    B, S = input_ids.shape
    e = p + "embeddings."
    x = (sd[e + "word_embeddings.weight"][input_ids]
         + sd[e + "token_type_embeddings.weight"][0]
         + sd[e + "position_embeddings.weight"][:S].unsqueeze(0))
    Hd = x.shape[-1]
    x = F.layer_norm(x, (Hd,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    bias = None
    if attention_mask is not None:
        bias = torch.zeros(B, 1, 1, S, dtype=x.dtype)
        bias.masked_fill_((attention_mask == 0).view(B, 1, 1, S), float("-inf"))
    d = Hd // heads
    i = 0
    while f"{p}encoder.layer.{i}.attention.self.query.weight" in sd:
        L = f"{p}encoder.layer.{i}."
        q = _linear(sd, L + "attention.self.query", x).view(B, S, heads, d).transpose(1, 2)
        k = _linear(sd, L + "attention.self.key", x).view(B, S, heads, d).transpose(1, 2)
        v = _linear(sd, L + "attention.self.value", x).view(B, S, heads, d).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(d)
        if bias is not None:
            scores = scores + bias
        ctx = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(B, S, Hd)
        a = _linear(sd, L + "attention.output.dense", ctx)
        x = F.layer_norm(a + x, (Hd,), sd[L + "attention.output.LayerNorm.weight"],
                         sd[L + "attention.output.LayerNorm.bias"], eps)
        f = F.gelu(_linear(sd, L + "intermediate.dense", x))  # exact erf GELU
        f = _linear(sd, L + "output.dense", f)
        x = F.layer_norm(f + x, (Hd,), sd[L + "output.LayerNorm.weight"],
                         sd[L + "output.LayerNorm.bias"], eps)
        i += 1
    return x


def text_encoder(sd: SD, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """src/text_encoder.py:95-127 -> [B,768] (CLS token, eval-mode dropout = identity)."""
    return bert_encoder(sd, input_ids, attention_mask)[:, 0, :]


# ---------------------------------------------------------------------------------- fusion + head
def _cross_attention(sd: SD, p: str, query: torch.Tensor, kv: torch.Tensor, heads: int
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """CrossModalAttention.forward (src/fusion_model.py:116-182) for [B,dim] inputs (seq len 1)."""
    B = query.shape[0]
    q = _linear(sd, p + "query_proj", query.unsqueeze(1))
    k = _linear(sd, p + "key_proj", kv.unsqueeze(1))
    v = _linear(sd, p + "value_proj", kv.unsqueeze(1))
    hidden = q.shape[-1]
    d = hidden // heads
    q = q.view(B, 1, heads, d).transpose(1, 2)
    k = k.view(B, 1, heads, d).transpose(1, 2)
    v = v.view(B, 1, heads, d).transpose(1, 2)
    w = torch.softmax((q @ k.transpose(-2, -1)) * d ** -0.5, dim=-1)  # [B,heads,1,1]
    out = (w @ v).transpose(1, 2).reshape(B, 1, hidden)
    return _linear(sd, p + "output_proj", out).squeeze(1), w


def attention_fusion(sd: SD, image_embedding: torch.Tensor, text_embedding: torch.Tensor,
                     heads: int = 8, use_residual: bool = True, eps: float = 1e-5,
                     p: str = "fusion.fusion_layer.") -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """src/fusion_model.py:245-291."""
    ip = _linear(sd, p + "image_proj", image_embedding)
    tp = _linear(sd, p + "text_proj", text_embedding)
    ia, w_i2t = _cross_attention(sd, p + "image_to_text_attention.", ip, tp, heads)
    ta, w_t2i = _cross_attention(sd, p + "text_to_image_attention.", tp, ip, heads)
    h = ip.shape[-1]
    io = F.layer_norm(ip + ia if use_residual else ia, (h,), sd[p + "layer_norm_image.weight"],
                      sd[p + "layer_norm_image.bias"], eps)
    to = F.layer_norm(tp + ta if use_residual else ta, (h,), sd[p + "layer_norm_text.weight"],
                      sd[p + "layer_norm_text.bias"], eps)
    fused = _linear(sd, p + "fusion.3", F.relu(_linear(sd, p + "fusion.0", torch.cat([io, to], -1))))
    return fused, {"image_to_text_attention": w_i2t, "text_to_image_attention": w_t2i}


def classification_head(sd: SD, x: torch.Tensor, p: str = "classifier.classifier.",
                        activation: str = "relu") -> torch.Tensor:
    """src/multimodal_classifier.py:73-83 (Linear, act, Dropout)* Linear; eval mode."""
    idx = sorted(int(k[len(p):].split(".")[0]) for k in sd if k.startswith(p) and k.endswith(".weight"))
    for j, i in enumerate(idx):
        x = _linear(sd, f"{p}{i}", x)
        if j + 1 < len(idx):
            x = F.gelu(x) if activation == "gelu" else F.relu(x)
    return x


def multimodal_forward(sd: SD, images: torch.Tensor, input_ids: torch.Tensor,
                       attention_mask: Optional[torch.Tensor], fusion_heads: int = 8
                       ) -> Dict[str, torch.Tensor]:
    """MultimodalClassifier.forward(return_embeddings=True) - src/multimodal_classifier.py:131-177."""
    with torch.no_grad():
        sd = {k: v.detach().float().cpu() for k, v in sd.items() if v.is_floating_point()}
        img = cnn_encoder(sd, images.float().cpu())
        txt = text_encoder(sd, input_ids.cpu(), None if attention_mask is None else attention_mask.cpu())
        fused, info = attention_fusion(sd, img, txt, heads=fusion_heads)
        logits = classification_head(sd, fused)
        return {"logits": logits, "probs": torch.softmax(logits, dim=-1), "image_embedding": img,
                "text_embedding": txt, "fused_embedding": fused, "attention_info": info}
