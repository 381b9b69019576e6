"""CNNEncoder - drop-in for the reference's src/cnn_encoder.py (ResNet50 branch).

Same constructor, attributes, parameter tree (state_dict keys `backbone.*`, `projection.{0,3}.*`)
and forward contract as src/cnn_encoder.py:14-242; the arithmetic of forward() runs in
libmrd_b200.so (tcgen05 implicit-GEMM convolutions with folded BN/ReLU/residual epilogues).
`backbone` is a torchvision ResNet held as the parameter container - exactly what the reference
holds - so checkpoints, freezing helpers and tree walks (src/train_multimodal.py:469) behave the same;
its Python forward is never executed.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ._module import B200Module
from .config import CNNEncoderConfig, get_config

_STEM_CHILDREN = ("conv1", "bn1", "relu", "maxpool")


class CNNEncoder(B200Module):
    _mrd_groups = {"": "cnn_encoder."}

    def __init__(self, config: Optional[CNNEncoderConfig] = None):
        super().__init__()
        config = get_config().cnn_encoder if config is None else config
        self.config = config
        self.backbone_name = config.backbone
        self.embedding_dim = config.embedding_dim
        self.backbone, feat = self._build_backbone()
        self.projection = nn.Sequential(
            nn.Linear(feat, config.embedding_dim),
            nn.ReLU(inplace=True),
            nn.Dropout(config.dropout),
            nn.Linear(config.embedding_dim, config.embedding_dim),
        )
        if config.freeze_backbone:
            self._freeze_backbone()
        elif config.freeze_layers > 0:
            self._freeze_layers(config.freeze_layers)

    # ---- construction ------------------------------------------------------------------------
    def _build_backbone(self) -> Tuple[nn.Module, int]:
        if self.backbone_name == "resnet50":
            from torchvision import models

            weights = models.ResNet50_Weights.IMAGENET1K_V2 if self.config.pretrained else None
            net = models.resnet50(weights=weights)
            feat = net.fc.in_features
            net.fc = nn.Identity()
            return net, feat
        if self.backbone_name == "efficientnet_b0":
            raise NotImplementedError(
                "efficientnet_b0 is outside the B200 hot path (SURVEY.md section 8: ResNet50 branch only)")
        raise ValueError(f"Unknown backbone: {self.backbone_name}")

    def _freeze_backbone(self) -> None:
        self.backbone.requires_grad_(False)

    def _freeze_layers(self, num_layers: int) -> None:
        # reference semantics (src/cnn_encoder.py:108-146): stem always, then layer1..layer{n}, n <= 4
        frozen = set(_STEM_CHILDREN) | {f"layer{i}" for i in range(1, min(num_layers, 4) + 1)}
        for name, child in self.backbone.named_children():
            if name in frozen:
                child.requires_grad_(False)

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,H,W] (H, W multiples of 32; 224 in the reference) -> [B, embedding_dim] fp32."""
        emb, _, _ = self._engine().cnn_encoder(x, self.embedding_dim)
        return emb

    def get_attention_layer(self) -> nn.Module:
        if self.backbone_name != "resnet50":
            raise ValueError(f"Unknown backbone: {self.backbone_name}")
        return self.backbone.layer4

    def get_intermediate_features(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(layer4 feature maps [B,2048,H/32,W/32], embedding) - src/cnn_encoder.py:200-226."""
        emb, _, fmap = self._engine().cnn_encoder(x, self.embedding_dim, want_map=True)
        return fmap, emb


class ResNet50Encoder(CNNEncoder):
    def __init__(self, embedding_dim: int = 512, pretrained: bool = True, dropout: float = 0.3,
                 freeze_layers: int = 0):
        super().__init__(CNNEncoderConfig(backbone="resnet50", embedding_dim=embedding_dim,
                                          pretrained=pretrained, dropout=dropout,
                                          freeze_layers=freeze_layers))


def create_cnn_encoder(backbone: str = "resnet50", embedding_dim: int = 512, pretrained: bool = True,
                       **kwargs) -> CNNEncoder:
    return CNNEncoder(CNNEncoderConfig(backbone=backbone, embedding_dim=embedding_dim,
                                       pretrained=pretrained, **kwargs))
