"""Batched prediction glue around the B200 forward.

The reference's `MultimodalPredictor` (src/predict.py:29-269) loads a checkpoint, turns PIL images / strings into
tensors (torchvision transforms, HF tokenizer - both stay reference code) and formats `model(...)["probs"]` into
per-sample dicts.  It works unchanged on the drop-in model (it only calls `model(images=..., input_ids=...,
attention_mask=...)`).  This module is the tensor-level half of it for large batches: host-resident tensors are
streamed through `MultimodalClassifier.forward_host` (H2D overlapped with compute) and the result rows are
formatted exactly like `MultimodalPredictor.predict_batch` (src/predict.py:199-269).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence

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
                          top_k: int = 3, micro_batch: int = 512) -> List[Dict]:
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
