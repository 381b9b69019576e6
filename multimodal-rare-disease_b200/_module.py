"""Shared base of the drop-in modules: lazily owns one Engine per device and keeps its packed
weights in step with the module's parameters."""

from __future__ import annotations

from typing import Dict, Iterator, Tuple

import torch
import torch.nn as nn

from .engine import Engine


class B200Module(nn.Module):
    # local prefix in this module's parameter tree -> canonical prefix the library expects
    _mrd_groups: Dict[str, str] = {}

    def _mrd_options(self) -> Dict[str, float]:
        return {}

    def _mrd_named(self) -> Iterator[Tuple[str, torch.Tensor]]:
        for local, canon in self._mrd_groups.items():
            sub = self.get_submodule(local.rstrip(".")) if local else self
            for n, p in sub.named_parameters():
                yield canon + n, p
            for n, b in sub.named_buffers():
                yield canon + n, b

    def _mrd_device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def _engine(self, allow_training: bool = False) -> Engine:
        """The engine for the device the parameters live on, with up-to-date packed weights."""
        if self.training and not allow_training:
            raise NotImplementedError(
                "this module's B200 path implements the inference forward (eval mode) only; call .eval() "
                "first.  Train mode is implemented for MultimodalClassifier.forward (the training step of "
                "SURVEY.md section 8(f)); there is deliberately no silent PyTorch fallback.")
        dev = self._mrd_device()
        eng = self.__dict__.get("_mrd_engine")
        if eng is None or eng.device != dev:
            eng = Engine(dev, self._mrd_options())
            self.__dict__["_mrd_engine"] = eng
        eng.sync_weights(self._mrd_named(), training=allow_training and self.training)
        return eng

    def configure_b200(self, img_chunk: int = 0, seq_chunk_tokens: int = 0, fp32_check=None) -> None:
        """Micro-batch sizes of the engine (images per ResNet pass, tokens per BERT pass).

        fp32_check=True switches this module's forwards to the library's plain-fp32 check kernels (no
        bf16 anywhere, BatchNorm un-folded): a slow verification mode that matches the reference's own
        fp32 forward to ~1e-5 (BASELINE.json north_star: 1e-4).  False switches back."""
        training, self.training = self.training, False
        try:
            eng = self._engine()
            eng.configure(img_chunk, seq_chunk_tokens)
            if fp32_check is not None:
                eng.set_option("fp32_check", 1.0 if fp32_check else 0.0)
        finally:
            self.training = training

    def __getstate__(self):
        # runtime-only state never travels with pickles / deep copies: the engine (a device context), the copy
        # stream and prefetched device tensors of forward_host, and the process group of data_parallel()
        d = self.__dict__.copy()
        for k in ("_mrd_engine", "_mrd_copy_stream", "_mrd_prefetched", "_mrd_ddp"):
            d.pop(k, None)
        return d
