"""MultimodalFusion - drop-in for the reference's src/fusion_model.py (fusion_type="attention").

Parameter tree identical to src/fusion_model.py:74-114,193-243,365-402 (`fusion_layer.image_proj`,
`.text_proj`, `.image_to_text_attention.{query,key,value,output}_proj`, `.text_to_image_attention.*`,
`.layer_norm_image`, `.layer_norm_text`, `.fusion.{0,3}`).  forward() follows src/fusion_model.py:245-291
in libmrd_b200.so.  Both modalities enter the cross attention with sequence length 1
(src/fusion_model.py:138-143), so the softmax is over one key: the returned weights are exactly 1
and attended = output_proj(value_proj(kv)); query_proj / key_proj stay in the state_dict but do not
influence the eval-mode output (they do in the reference either).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ._module import B200Module
from .config import FusionConfig, get_config


class CrossModalAttention(nn.Module):
    """Parameter container with the reference's names (src/fusion_model.py:81-114)."""

    def __init__(self, query_dim: int, key_dim: int, hidden_dim: int, num_heads: int = 8,
                 dropout: float = 0.1):
        super().__init__()
        assert hidden_dim % num_heads == 0, "hidden_dim must be divisible by num_heads"
        self.num_heads, self.hidden_dim = num_heads, hidden_dim
        self.head_dim = hidden_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.query_proj = nn.Linear(query_dim, hidden_dim)
        self.key_proj = nn.Linear(key_dim, hidden_dim)
        self.value_proj = nn.Linear(key_dim, hidden_dim)
        self.output_proj = nn.Linear(hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)


class AttentionFusion(nn.Module):
    """Parameter container with the reference's names (src/fusion_model.py:193-243)."""

    def __init__(self, config: Optional[FusionConfig] = None):
        super().__init__()
        config = get_config().fusion if config is None else config
        self.config = config
        h = self.hidden_dim = config.hidden_dim
        self.image_proj = nn.Linear(config.image_proj_dim, h)
        self.text_proj = nn.Linear(config.text_proj_dim, h)
        self.image_to_text_attention = CrossModalAttention(h, h, h, config.num_attention_heads,
                                                           config.dropout)
        self.text_to_image_attention = CrossModalAttention(h, h, h, config.num_attention_heads,
                                                           config.dropout)
        self.layer_norm_image = nn.LayerNorm(h)
        self.layer_norm_text = nn.LayerNorm(h)
        self.fusion = nn.Sequential(nn.Linear(2 * h, h), nn.ReLU(inplace=True),
                                    nn.Dropout(config.dropout), nn.Linear(h, h))
        self.use_residual = config.use_residual


class MultimodalFusion(B200Module):
    _mrd_groups = {"": "fusion."}

    def __init__(self, config: Optional[FusionConfig] = None):
        super().__init__()
        config = get_config().fusion if config is None else config
        self.config = config
        self.fusion_type = config.fusion_type
        if config.fusion_type == "attention":
            self.fusion_layer = AttentionFusion(config)
        elif config.fusion_type in ("concatenation", "gated"):
            raise NotImplementedError(
                f"fusion_type={config.fusion_type!r} is outside the B200 hot path "
                "(SURVEY.md section 8: the default attention fusion only)")
        else:
            raise ValueError(f"Unknown fusion type: {config.fusion_type}")

    def _mrd_options(self):
        return {"fusion_heads": self.config.num_attention_heads,
                "fusion_residual": 1.0 if self.config.use_residual else 0.0,
                "fusion_ln_eps": self.fusion_layer.layer_norm_image.eps}

    def forward(self, image_embedding: torch.Tensor, text_embedding: torch.Tensor
                ) -> Tuple[torch.Tensor, Optional[dict]]:
        fused, a_i2t, a_t2i = self._engine().fusion(image_embedding, text_embedding,
                                                    self.config.hidden_dim,
                                                    self.config.num_attention_heads)
        return fused, {"image_to_text_attention": a_i2t, "text_to_image_attention": a_t2i}


def create_fusion_module(fusion_type: str = "attention", image_dim: int = 512, text_dim: int = 768,
                         hidden_dim: int = 512, **kwargs) -> MultimodalFusion:
    return MultimodalFusion(FusionConfig(fusion_type=fusion_type, image_proj_dim=image_dim,
                                         text_proj_dim=text_dim, hidden_dim=hidden_dim, **kwargs))
