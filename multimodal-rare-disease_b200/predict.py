"""Batched prediction glue around the B200 forward.

The reference's `MultimodalPredictor` (src/predict.py:29-269) loads a checkpoint, turns PIL images / strings into
tensors (torchvision transforms, HF tokenizer - both stay reference code) and formats `model(...)["probs"]` into
per-sample dicts.  It works unchanged on the drop-in model (it only calls `model(images=..., input_ids=...,
attention_mask=...)`).  This module is the tensor-level half of it for large batches: host-resident tensors are
streamed through `MultimodalClassifier.forward_host` (H2D overlapped with compute) and the result rows are
formatted exactly like `MultimodalPredictor.predict_batch` (src/predict.py:199-269).
"""

from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch


def format_predictions(probs: torch.Tensor, class_names: Optional[Sequence[str]] = None, top_k: int = 3,
                       first_index: int = 0) -> List[Dict]:
    """[B, C] probabilities -> the list of dicts `predict_batch` returns (src/predict.py:241-267): per sample the
    top_k classes by probability (ties broken like numpy's argsort()[::-1]: the higher class index first)."""
    if probs.dim() != 2:
        raise ValueError(f"probs must be [B, C], got {tuple(probs.shape)}")
    names = list(class_names) if class_names is not None else []
    p = probs.detach().float().cpu()
    B, Cn = p.shape
    k = max(0, min(int(top_k), Cn))
    out = []
    if k:
        # stable sort of the reversed class order == numpy argsort()[::-1] on ties
        order = torch.argsort(p.flip(-1), dim=-1, descending=True, stable=True)[:, :k]
        order = (Cn - 1) - order
    for i in range(B):
        preds = []
        for j in range(k):
            idx = int(order[i, j])
            preds.append({"syndrome": names[idx] if idx < len(names) else f"Class_{idx}", "class_id": idx,
                          "confidence": float(p[i, idx])})
        out.append({"sample_idx": first_index + i, "predictions": preds,
                    "top_prediction": preds[0] if preds else None})
    return out


@torch.no_grad()
def predict_batch_tensors(model, images: torch.Tensor, input_ids: torch.Tensor,
                          attention_mask: Optional[torch.Tensor], class_names: Optional[Sequence[str]] = None,
                          top_k: int = 3, micro_batch: int = 2048) -> List[Dict]:
    """Preprocessed tensors (on the host or on the model's device) -> `predict_batch`-style results.
    Host tensors should be pinned for the copies to overlap the kernels."""
    if images.shape[0] != input_ids.shape[0]:
        raise ValueError("Number of images must match number of texts")   # src/predict.py:214
    # model.train(flag) recurses: a model trained as `model.train(); model.cnn_encoder.backbone.eval()` must come
    # back with exactly the per-module flags it had (or the next step would switch BatchNorm to batch statistics)
    modes = [(m, m.training) for m in model.modules()]
    model.eval()
    try:
        if images.device.type == "cpu":
            out = model.forward_host(images, input_ids, attention_mask, micro_batch=micro_batch)
        else:
            out = model(images=images, input_ids=input_ids, attention_mask=attention_mask)
        return format_predictions(out["probs"], class_names, top_k)
    finally:
        for m, flag in modes:
            m.training = flag


def load_checkpoint(model, checkpoint_path: Union[str, Path], strict: bool = True) -> Dict:
    """`MultimodalPredictor._load_checkpoint` (src/predict.py:73-82): reads a checkpoint written by the reference's
    trainers (src/train.py:394-437: {"model_state_dict": ..., "optimizer_state_dict": ..., ...}) - or a bare
    state_dict - into the drop-in model.  The state_dict keys are the reference's own, so nothing is renamed; the
    library's packed bf16 weights are rebuilt on the next forward (the parameters' version counters changed).
    Returns the checkpoint dict (epoch, metrics, optimizer state for a caller that resumes training)."""
    path = Path(checkpoint_path)
    if not path.exists():
        raise FileNotFoundError(f"Checkpoint not found: {path}")     # src/predict.py:77-78
    ckpt = torch.load(path, map_location="cpu")
    sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    model.load_state_dict(sd, strict=strict)
    return ckpt if isinstance(ckpt, dict) else {"model_state_dict": sd}


@torch.no_grad()
def collect_predictions(model, loader: Iterable, mode: str = "multimodal", device=None,
                        micro_batch: int = 2048) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """`Evaluator.collect_predictions` (src/evaluate.py:79-123) on the B200 path: walks a loader of the reference's
    batch dicts ({"image", "input_ids", "attention_mask", "label"}; image_only loaders may yield (image, label)
    tuples) and returns (predictions i64 [N], true_labels i64 [N], probabilities f32 [N,C]) on the host - the three
    arrays the reference's compute_metrics / confusion-matrix code consumes.
    Host batches of the multimodal mode go through `forward_host` (H2D of micro-batch i+1 under the kernels of
    micro-batch i); argmax runs on the device and one transfer per batch brings predictions and probabilities back."""
    if mode not in ("multimodal", "image_only", "text_only"):
        raise ValueError(f"Unknown mode: {mode}")
    modes = [(m, m.training) for m in model.modules()]
    model.eval()
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    preds, labels, probs = [], [], []
    try:
        for batch in loader:
            if mode == "multimodal":
                images, ids, mask, y = batch["image"], batch["input_ids"], batch["attention_mask"], batch["label"]
                if images.device.type == "cpu" and hasattr(model, "forward_host"):
                    out = model.forward_host(images, ids, mask, micro_batch=micro_batch)
                else:
                    out = model(images.to(dev), ids.to(dev), mask.to(dev))
            elif mode == "image_only":
                images, y = (batch[0], batch[1]) if isinstance(batch, (tuple, list)) else (batch["image"], batch["label"])
                out = model(images.to(dev))
            else:
                y = batch["label"]
                out = model(batch["input_ids"].to(dev), batch["attention_mask"].to(dev))
            p = out["probs"]
            both = torch.cat([out["logits"].argmax(dim=-1, keepdim=True).to(p.dtype), p], dim=1).cpu()
            preds.append(both[:, 0].long())
            probs.append(both[:, 1:])
            labels.append(y.detach().cpu().long() if torch.is_tensor(y) else torch.as_tensor(list(y), dtype=torch.long))
    finally:
        for m, flag in modes:
            m.training = flag
    if not preds:
        return torch.empty(0, dtype=torch.long), torch.empty(0, dtype=torch.long), torch.empty(0, 0)
    return torch.cat(preds), torch.cat(labels), torch.cat(probs)
