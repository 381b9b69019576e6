"""FusedAdamW - the optimizer step of the reference's training loop on the B200 path.

The reference does `nn.utils.clip_grad_norm_(model.parameters(), 1.0); optimizer.step()` with
`torch.optim.AdamW` (src/train.py:193-198,307-320; per-group learning rates in
src/train_multimodal.py:422-454).  That keeps working unchanged with the drop-in model; this class is the
optional native replacement: gradient-norm reduction, clipping and the AdamW update of every parameter in two
launches of libmrd_b200.so (mrd_adamw_step), with no host synchronisation.  Same math, hyper-parameters, param-group
semantics and state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) as torch.optim.AdamW(amsgrad=False), so
`state_dict()` / `load_state_dict()` are interchangeable with it (checkpoints of src/train.py:394-437).
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

_TENSOR = np.dtype([("p", np.uint64), ("g", np.uint64), ("m", np.uint64), ("v", np.uint64), ("n", np.int64),
                    ("lr", np.float32), ("wd", np.float32)])
_CHUNK = 1 << 16


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        """max_grad_norm: clip the global gradient norm inside step() (replaces the separate clip_grad_norm_ call;
        the .grad tensors themselves are left unscaled).  None / 0 = no clipping."""
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self._lib = None
        self._plan = None
        self.last_grad_norm = None   # device scalar (fp32), the unclipped global norm of the last step

    # ------------------------------------------------------------------ tables
    def _params(self):
        return [(p, g) for g in self.param_groups for p in g["params"]]

    def _build(self, device):
        items = self._params()
        betas = {tuple(g["betas"]) for g in self.param_groups}
        epss = {g["eps"] for g in self.param_groups}
        if len(betas) != 1 or len(epss) != 1:
            raise NotImplementedError("FusedAdamW: betas and eps must be the same in every parameter group")
        chunk_t, chunk_o = [], []
        for i, (p, _) in enumerate(items):
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.MrdError("FusedAdamW: parameters must be contiguous fp32 tensors on one CUDA device")
            st = self.state[p]
            if not st:
                st["step"] = torch.zeros((), dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            for off in range(0, p.numel(), _CHUNK):
                chunk_t.append(i)
                chunk_o.append(off)
        # two pinned staging tables used alternately: the host may run a full step ahead of the device, so a
        # table is rewritten only after the copy that read it has completed (event below)
        hosts = [torch.empty(len(items) * _TENSOR.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._plan = {
            "device": device, "n": len(items), "key": tuple(id(p) for p, _ in items),
            "hosts": hosts, "recs": [h.numpy().view(_TENSOR) for h in hosts], "events": [None, None], "turn": 0,
            "dev": torch.empty(len(items) * _TENSOR.itemsize, dtype=torch.uint8, device=device),
            "chunk_t": torch.tensor(chunk_t, dtype=torch.int32, device=device),
            "chunk_o": torch.tensor(chunk_o, dtype=torch.int64, device=device),
            "sq": torch.zeros(1, dtype=torch.float32, device=device),
        }

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = self._params()
        if not items:
            return loss
        device = items[0][0].device
        if device.type != "cuda":
            raise _lib.MrdError("FusedAdamW runs on CUDA parameters only (use torch.optim.AdamW on CPU)")
        if self._lib is None:
            self._lib = _lib.load()
        if self._plan is None or self._plan["key"] != tuple(id(p) for p, _ in items) or self._plan["device"] != device:
            self._build(device)
        pl = self._plan
        turn = pl["turn"]
        pl["turn"] = turn ^ 1
        if pl["events"][turn] is not None:
            pl["events"][turn].synchronize()
        rec = pl["recs"][turn]
        steps = set()
        any_grad = False
        for i, (p, g) in enumerate(items):
            st = self.state[p]
            grad = p.grad
            if grad is not None:
                if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.is_sparse:
                    raise _lib.MrdError("FusedAdamW: gradients must be dense contiguous fp32")
                st["step"] += 1
                steps.add(int(st["step"]))
                any_grad = True
            rec[i] = (p.data_ptr(), 0 if grad is None else grad.data_ptr(), st["exp_avg"].data_ptr(),
                      st["exp_avg_sq"].data_ptr(), p.numel(), g["lr"], g["weight_decay"])
        if not any_grad:
            return loss
        if len(steps) != 1:
            raise NotImplementedError("FusedAdamW: parameters with different step counts (a parameter that got its "
                                      "first gradient later than the others) are not supported")
        pl["dev"].copy_(pl["hosts"][turn], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        pl["events"][turn] = ev
        b1, b2 = self.param_groups[0]["betas"]
        with torch.cuda.device(device):
            _lib.check(self._lib.mrd_adamw_step(
                pl["dev"].data_ptr(), pl["n"], pl["chunk_t"].data_ptr(), pl["chunk_o"].data_ptr(),
                pl["chunk_t"].numel(), _CHUNK, float(b1), float(b2), float(self.param_groups[0]["eps"]),
                steps.pop(), self.max_grad_norm, pl["sq"].data_ptr(),
                C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "mrd_adamw_step")
        # the kernel wrote the parameters through raw pointers: bump their torch version counters so that everything
        # keyed on `_version` sees the update (Engine.sync_weights re-packs the library's bf16 copies of a group only
        # when a (data_ptr, _version) signature changed; autograd's saved-tensor checks rely on it as well)
        torch.autograd.graph.increment_version([p for p, _ in items if p.grad is not None])
        self.last_grad_norm = pl["sq"].sqrt()
        return loss
