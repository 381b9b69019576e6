"""Configuration dataclasses of the hot path.

Field-for-field mirrors of the model-defining dataclasses in the reference's src/config.py
(CNNEncoderConfig :58-68, TextEncoderConfig :70-80, FusionConfig :83-94, ClassifierConfig :97-105,
Config :182-217).  The reference's own dataclass instances are accepted everywhere as well (the
modules only read attributes).  Data/training/evaluation settings are not part of this path; unlike
the reference's Config, constructing this one has no filesystem side effects.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


SYNDROME_NAMES = ("Cornelia de Lange Syndrome", "Williams-Beuren Syndrome", "Noonan Syndrome", "Kabuki Syndrome",
                  "KBG Syndrome", "Angelman Syndrome", "Rubinstein-Taybi Syndrome", "Smith-Magenis Syndrome",
                  "Nicolaides-Baraitser Syndrome", "22q11.2 Deletion Syndrome")


@dataclass
class CNNEncoderConfig:
    backbone: str = "resnet50"            # the B200 path implements ResNet50 only (efficientnet_b0 raises)
    pretrained: bool = True               # ImageNet weights through torchvision, as the reference does
    embedding_dim: int = 512
    freeze_backbone: bool = True          # the training step requires the backbone frozen (reference default)
    freeze_layers: int = 6                # only read when freeze_backbone is False
    dropout: float = 0.5


@dataclass
class TextEncoderConfig:
    model_name: str = "dmis-lab/biobert-base-cased-v1.2"   # any BERT-base shaped checkpoint
    embedding_dim: int = 768
    max_length: int = 128                 # tokenizer side; the kernels take S <= 512
    freeze_embeddings: bool = False       # frozen parameters simply get no gradient slot
    freeze_layers: int = 0
    dropout: float = 0.1
    use_pooler_output: bool = False       # True is outside the path (CLS row of the last layer is used)


@dataclass
class FusionConfig:
    fusion_type: str = "attention"        # concatenation / gated are outside the path
    hidden_dim: int = 512
    num_attention_heads: int = 8          # heads of the two length-1 cross attentions
    dropout: float = 0.3
    use_residual: bool = True             # LayerNorm(proj + attended) vs LayerNorm(attended)
    image_proj_dim: int = 512             # = CNNEncoderConfig.embedding_dim
    text_proj_dim: int = 768              # = BERT hidden size


@dataclass
class ClassifierConfig:
    hidden_dims: List[int] = field(default_factory=lambda: [256, 128])
    num_classes: int = 10                 # the final GEMV + softmax kernel handles up to 32
    dropout: float = 0.5
    activation: str = "relu"              # relu / gelu in eval mode, relu in the training step


@dataclass
class Config:
    cnn_encoder: CNNEncoderConfig = field(default_factory=CNNEncoderConfig)
    text_encoder: TextEncoderConfig = field(default_factory=TextEncoderConfig)
    fusion: FusionConfig = field(default_factory=FusionConfig)
    classifier: ClassifierConfig = field(default_factory=ClassifierConfig)
    # class index -> syndrome label, the order the reference's classification head is trained with
    syndrome_names: List[str] = field(default_factory=lambda: list(SYNDROME_NAMES))
    seed: int = 42                        # kept for the callers; nothing here draws random numbers


config = Config()


def get_config() -> Config:
    """Module-level default configuration (reference: src/config.py:221-226)."""
    return config


# BioBERT-base-cased-v1.2 architecture (what AutoConfig.from_pretrained returns for the reference's
# default model_name); used for random-init construction when the checkpoint cannot be fetched.
BIOBERT_BASE = dict(
    vocab_size=28996,
    hidden_size=768,
    num_hidden_layers=12,
    num_attention_heads=12,
    intermediate_size=3072,
    hidden_act="gelu",
    hidden_dropout_prob=0.1,
    attention_probs_dropout_prob=0.1,
    max_position_embeddings=512,
    type_vocab_size=2,
    initializer_range=0.02,
    layer_norm_eps=1e-12,
    pad_token_id=0,
)
