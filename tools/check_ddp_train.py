#!/usr/bin/env python
"""Data-parallel training check on real GPUs (run under torchrun with 2+ ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_train.py

Every rank runs one training step on its shard of a global batch with model.data_parallel() (gradient bucket
averaged over NCCL inside loss.backward()); rank 0 additionally runs the whole batch alone.  With equal shards
the averaged per-rank gradients of the mean loss equal the single-process gradients up to fp32 summation order
(the frozen backbone is in eval mode so no cross-sample BatchNorm statistics enter).  Prints one JSON line."""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F

    import mrd_b200
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    per, S = 4, 64
    total = per * world
    images, ids, mask = synth.make_inputs(total, S, 91, [S - 3 * (i % 7) for i in range(total)], H=64, W=64)
    labels = torch.arange(total) % 10

    def build():
        m = synth.build_model(0)
        m.load_state_dict(synth.train_weights(0))
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
        mc = m.text_encoder.model_config
        mc.hidden_dropout_prob = mc.attention_probs_dropout_prob = 0.0
        m = m.to(dev).train()
        m.cnn_encoder.backbone.eval()
        return m

    model = build().data_parallel(True)
    lo, hi = mrd_b200.shard_bounds(total, world, rank)
    out = model(images[lo:hi].to(dev), ids[lo:hi].to(dev), mask[lo:hi].to(dev))
    F.cross_entropy(out["logits"], labels[lo:hi].to(dev)).backward()
    torch.cuda.synchronize()
    mine = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    # every rank must hold the same averaged gradients
    probe = torch.stack([g.double().sum() for g in mine.values()])
    gathered = [torch.empty_like(probe) for _ in range(world)]
    dist.all_gather(gathered, probe)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    res = None
    if rank == 0:
        single = build()
        out = single(images.to(dev), ids.to(dev), mask.to(dev))
        F.cross_entropy(out["logits"], labels.to(dev)).backward()
        torch.cuda.synchronize()
        num = den = 0.0
        worst = (0.0, "")
        for k, p in single.named_parameters():
            if p.grad is None:
                continue
            d = (mine[k] - p.grad).double().pow(2).sum().item()
            n = p.grad.double().pow(2).sum().item()
            num += d
            den += n
            if n > 1e-20 and (d / n) ** 0.5 > worst[0]:
                worst = ((d / n) ** 0.5, k)
        res = {"world": world, "ranks_agree_bitwise": bool(same), "global_rel_l2_vs_single_process": (num / den) ** 0.5,
               "worst_tensor": worst, "n_grads": len(mine)}
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not (res["ranks_agree_bitwise"] and res["global_rel_l2_vs_single_process"] < 2e-2):
        sys.exit(1)


if __name__ == "__main__":
    main()
