"""Batch-sharded data parallelism for the forward: one process per GPU, replicated weights.

Every sample is independent in eval mode (BatchNorm uses running statistics and no op mixes samples,
src/multimodal_classifier.py:157-167), so the only exchange is the gather of the [B,10] logits.  The
head kernel writes this rank's logits straight into its slot of the gather buffer and a single
in-place all_gather_into_tensor (NCCL over NVLink/NVSwitch; gloo in the CPU tests) fills the rest - no
staging copy, no other collective.  The reference has no multi-device code.
"""

from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` samples owned by `rank`; sizes differ by at most one
    and the larger shards come first."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallelForward:
    """Runs `local_forward(images, ids, mask, logits_out)` on this rank's shard and gathers logits.

    local_forward must write float32 logits [n_local, num_classes] into `logits_out` (it may also
    return a dict of extra per-shard outputs).  The wrapper is backend-agnostic so the host logic is
    testable with gloo on CPU; on GPUs `local_forward` is MultimodalClassifier.forward(...,
    logits_out=...).
    """

    def __init__(self, local_forward: Callable, num_classes: int, group: Optional[dist.ProcessGroup] = None):
        self.local_forward = local_forward
        self.num_classes = num_classes
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._buf = None

    def _buffer(self, total: int, device: torch.device) -> torch.Tensor:
        per = -(-total // self.world)  # equal slots so one all_gather_into_tensor suffices
        shape = (self.world, per, self.num_classes)
        if self._buf is None or self._buf.shape != shape or self._buf.device != device:
            self._buf = torch.zeros(shape, dtype=torch.float32, device=device)
        return self._buf

    def forward_shard(self, images, input_ids, attention_mask, total: int) -> torch.Tensor:
        """Inputs are THIS RANK's shard (shard_bounds(total, world, rank)).  Returns logits
        [total, num_classes] on every rank."""
        lo, hi = shard_bounds(total, self.world, self.rank)
        n = hi - lo
        if images.shape[0] != n:
            raise ValueError(f"rank {self.rank} expects {n} samples, got {images.shape[0]}")
        buf = self._buffer(total, images.device)
        slot = buf[self.rank]
        self.local_forward(images, input_ids, attention_mask, slot[:n])
        if self.world > 1:
            dist.all_gather_into_tensor(buf.view(-1), slot.reshape(-1), group=self.group)
        per = buf.shape[1]
        if total == per * self.world:
            return buf.view(total, self.num_classes)
        parts = []
        for r in range(self.world):
            rlo, rhi = shard_bounds(total, self.world, r)
            parts.append(buf[r, : rhi - rlo])
        return torch.cat(parts, 0)

    def forward_shard_host(self, host_forward: Callable, images, input_ids, attention_mask, total: int,
                           device: torch.device) -> torch.Tensor:
        """forward_shard for a shard that still lives in host memory: host_forward(images, ids, mask,
        logits_out) streams it to the device in micro-batches (MultimodalClassifier.forward_host)."""
        lo, hi = shard_bounds(total, self.world, self.rank)
        n = hi - lo
        if images.shape[0] != n:
            raise ValueError(f"rank {self.rank} expects {n} samples, got {images.shape[0]}")
        buf = self._buffer(total, device)
        slot = buf[self.rank]
        host_forward(images, input_ids, attention_mask, slot[:n])
        if self.world > 1:
            dist.all_gather_into_tensor(buf.view(-1), slot.reshape(-1), group=self.group)
        if total == buf.shape[1] * self.world:
            return buf.view(total, self.num_classes)
        return torch.cat([buf[r, : shard_bounds(total, self.world, r)[1] - shard_bounds(total, self.world, r)[0]]
                          for r in range(self.world)], 0)

    def forward_global(self, images, input_ids, attention_mask) -> torch.Tensor:
        """Inputs are the full batch (replicated on every rank); each rank computes its shard."""
        total = images.shape[0]
        lo, hi = shard_bounds(total, self.world, self.rank)
        mask = attention_mask[lo:hi] if attention_mask is not None else None
        return self.forward_shard(images[lo:hi], input_ids[lo:hi], mask, total)


def allreduce_mean_(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Averages a gradient bucket over the data-parallel ranks in place (what DistributedDataParallel
    does; the reference has no multi-device training).  The B200 training step
    writes every parameter gradient into ONE flat fp32 buffer (Engine.train_backward), so the whole
    exchange is a single all-reduce: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests."""
    if not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    return flat


def allreduce_mean_async(bucket: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Starts the averaging all-reduce of one gradient bucket and returns a handle for wait_allreduce (None
    without a process group).  The training backward hands over its flat gradient buffer bucket by bucket while
    the remaining layers are still being differentiated (bucketed, overlapped)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    work = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return (work, bucket, dist.get_world_size(group))


def wait_allreduce(handle) -> None:
    if handle is None:
        return
    work, bucket, world = handle
    work.wait()
    bucket.mul_(1.0 / world)


def rank_seed(seed: int, group: Optional[dist.ProcessGroup] = None) -> int:
    """Per-rank dropout seed for data-parallel training: the step seed mixed with the rank (rank 0 keeps the
    single-process seed), so shards do not share dropout masks when every rank seeds torch identically."""
    if not dist.is_initialized():
        return seed
    rank = dist.get_rank(group)
    return (seed ^ (rank * 0x9E3779B97F4A7C15)) & 0x3FFFFFFFFFFFFFFF


def broadcast_buffers_(buffers, group: Optional[dist.ProcessGroup] = None, src: int = 0) -> None:
    """Overwrite `buffers` (BatchNorm running statistics) on every rank with rank `src`'s values: one flat
    broadcast, as DistributedDataParallel does with broadcast_buffers=True."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or not buffers:
        return
    flat = torch.cat([b.detach().reshape(-1).float() for b in buffers])
    dist.broadcast(flat, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    off = 0
    with torch.no_grad():
        for b in buffers:
            n = b.numel()
            b.copy_(flat[off:off + n].view_as(b))
            off += n
