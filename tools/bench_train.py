#!/usr/bin/env python
"""BASELINE.json configs[4]: multimodal TRAINING step (forward + backward + AdamW), batch 64 per GPU,
224x224 images, 128 tokens with padding masks L ~ U{16..128}, default freeze configuration (ResNet50
backbone frozen, its BatchNorm on batch statistics as a bare model.train() gives), gradient all-reduce over
NCCL at N > 1.  The step is the reference's own loop body (src/train.py:247-321): zero_grad / forward /
CrossEntropyLoss / backward / clip_grad_norm_(1.0) / AdamW(lr 5e-5, wd 0.05).step().

    python tools/bench_train.py [--steps 20 --warmup 5 --per-gpu-batch 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py

Prints one JSON line (rank 0): samples/s over all ranks, ms per step, and - from one extra profiled step -
the device time per kernel label.  Not the repo's bench line (that is bench.py, the inference forward)."""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--per-gpu-batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--bn-eval", action="store_true", help="frozen backbone in eval mode (running statistics)")
    ap.add_argument("--fused-adamw", action="store_true", help="torch.optim.AdamW(fused=True)")
    ap.add_argument("--native-adamw", action="store_true",
                    help="mrd_b200.FusedAdamW(max_grad_norm=1): clip + AdamW in two library launches")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import torch.nn as nn

    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.per_gpu_batch, args.seq
    model = synth.build_model(0).to(dev)       # same seed on every rank = identical replicas
    model.train()
    if args.bn_eval:
        model.cnn_encoder.backbone.eval()
    model.data_parallel(world > 1)
    g = torch.Generator().manual_seed(1234 + rank)
    images = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    ids = torch.randint(1, 28996, (B, S), generator=g)
    lengths = torch.randint(16, S + 1, (B,), generator=g)
    mask = (torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1)).long()
    ids = ids * mask
    ids[:, 0] = 101
    ids, mask = ids.to(dev), mask.to(dev)
    labels = torch.randint(0, 10, (B,), generator=g).to(dev)
    if args.native_adamw:
        import mrd_b200

        opt = mrd_b200.FusedAdamW(model.parameters(), lr=5e-5, weight_decay=0.05, max_grad_norm=1.0)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.05, fused=args.fused_adamw or None)
    crit = nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(images, ids, mask)["logits"], labels)
        loss.backward()
        if not args.native_adamw:
            nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(step().item())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng = model._engine(allow_training=True)
    l0 = eng.launch_count
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_step = ms / args.steps
    line = None
    if rank == 0:
        # one profiled step: device time per label of the LIBRARY's kernels (the optimizer and the loss are torch's).
        # Rank 0 runs it alone, so the gradient all-reduce must be off for it (the other ranks are not in this step).
        model.data_parallel(False)
        eng.profile(True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        opt.zero_grad(set_to_none=True)
        t0.record()
        out = model(images, ids, mask)["logits"]
        t1.record()
        l = crit(out, labels)
        l.backward()
        rows = eng.profile_report()
        eng.profile(False)
        tot = sum(r["ms"] for r in rows)
        prof = [{"label": r["label"], "ms": round(r["ms"], 4), "launches": r["launches"], "share": round(r["ms"] / tot, 4)}
                for r in sorted(rows, key=lambda r: -r["ms"])]
        if args.profile_out:
            with open(args.profile_out, "w") as fh:
                fh.write("label,launches,total_ms,share\n")
                for r in prof:
                    fh.write(f'{r["label"]},{r["launches"]},{r["ms"]},{r["share"]}\n')
        live = float(((mask != 0) | (torch.arange(S, device=dev) == 0)).float().mean().item())
        line = {"metric": "multimodal training step samples/s (fwd+bwd+AdamW)", "value": world * B / (ms_step * 1e-3),
                "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "per_gpu_batch": B, "seq_len": S, "dtype": "bf16 (fp32 master weights, fp32 grads)",
                "bn": "running statistics" if args.bn_eval else "batch statistics",
                "optimizer": "mrd_b200.FusedAdamW (clip fused)" if args.native_adamw else
                             ("torch AdamW fused" if args.fused_adamw else "torch AdamW (foreach) + clip_grad_norm_"),
                "library_kernel_ms_profiled_step": round(tot, 3), "library_launches_per_step": launches // args.steps,
                "loss_first_last": [losses[0] if losses else None, loss.item()], "live_token_fraction": live,
                "trainable_params": sum(p.numel() for p in model.parameters() if p.requires_grad),
                "top_kernels": prof[:14]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
