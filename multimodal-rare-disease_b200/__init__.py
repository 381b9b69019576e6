"""B200-native drop-in for the hot path of ArshvirSk/Multimodal-Rare-Disease: the batched
MultimodalClassifier forward (ResNet50 CNNEncoder + BioBERT-base TextEncoder + attention fusion +
classification head), computed by hand-written sm_100a kernels behind a C ABI (libmrd_b200.so).

The directory name contains a hyphen; import it with
    importlib.import_module("multimodal-rare-disease_b200")
or through the `mrd_b200` alias module at the repository root.
"""

from .config import (BIOBERT_BASE, ClassifierConfig, CNNEncoderConfig, Config, FusionConfig,
                     TextEncoderConfig, get_config)
from .cnn_encoder import CNNEncoder, ResNet50Encoder, create_cnn_encoder
from .text_encoder import BioBERTEncoder, TextEncoder, create_text_encoder
from .fusion_model import AttentionFusion, CrossModalAttention, MultimodalFusion, create_fusion_module
from .multimodal_classifier import (ClassificationHead, ImageOnlyClassifier, MultimodalClassifier,
                                    TextOnlyClassifier, create_baseline_classifiers,
                                    create_multimodal_classifier)
from .optim import FusedAdamW
from .predict import collect_predictions, format_predictions, load_checkpoint, predict_batch_tensors
from .parallel import DataParallelForward, allreduce_mean_, shard_bounds
from ._lib import MrdError

__version__ = "0.1.0"
