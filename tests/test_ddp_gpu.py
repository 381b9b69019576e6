"""Data-parallel paths on real GPUs (needs >= 2 visible B200s: `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py
-m gpu`; skipped on a single-GPU box).  One process per GPU over NCCL, as bench.py and the training loop run.

  * training: every rank runs one step on its shard with model.data_parallel() - the gradient buckets (one per
    backward stage) are all-reduced while the remaining stages run.  All ranks must end with BIT-IDENTICAL
    gradients, equal to the single-process gradients of the whole batch up to fp32 summation order, and after an
    optimizer step with bit-identical parameters;
  * inference: the gathered logits of the sharded forward are bit-identical to the single-process forward of the
    whole batch (no op of the forward mixes samples: SURVEY.md section 4(iv)).
"""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _build(dev, train):
    import synth

    m = synth.build_model(0)
    if train:
        m.load_state_dict(synth.train_weights(0))
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
        mc = m.text_encoder.model_config
        mc.hidden_dropout_prob = mc.attention_probs_dropout_prob = 0.0
    m = m.to(dev)
    if train:
        m.train()
        m.cnn_encoder.backbone.eval()
    return m


def _worker(rank, world, port, q):
    import mrd_b200
    import synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        per, S = 4, 64
        total = per * world
        images, ids, mask = synth.make_inputs(total, S, 91, [S - 3 * (i % 7) for i in range(total)], H=64, W=64)
        labels = torch.arange(total) % 10
        lo, hi = mrd_b200.shard_bounds(total, world, rank)
        # ---- training step, bucketed + overlapped gradient all-reduce
        model = _build(dev, True).data_parallel(True)
        opt = mrd_b200.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
        out = model(images[lo:hi].to(dev), ids[lo:hi].to(dev), mask[lo:hi].to(dev))
        F.cross_entropy(out["logits"], labels[lo:hi].to(dev)).backward()
        torch.cuda.synchronize()
        mine = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        opt.step()
        torch.cuda.synchronize()
        digest = synth.tensor_digest(torch.cat([g.flatten() for g in mine.values()]))
        pdigest = synth.tensor_digest(torch.cat([p.detach().flatten() for p in model.parameters()]))
        # ---- sharded inference forward + logits gather
        emodel = _build(dev, False)
        dp = mrd_b200.DataParallelForward(lambda im, i, m, o: emodel(im, i, m, logits_out=o), emodel.num_classes)
        with torch.no_grad():
            gathered = dp.forward_shard(images[lo:hi].to(dev), ids[lo:hi].to(dev), mask[lo:hi].to(dev), total)
        res = {"rank": rank, "grad_digest": digest, "param_digest": pdigest,
               "logits_digest": synth.tensor_digest(gathered)}
        if rank == 0:
            single = _build(dev, True)
            o1 = single(images.to(dev), ids.to(dev), mask.to(dev))
            F.cross_entropy(o1["logits"], labels.to(dev)).backward()
            with torch.no_grad():
                whole = emodel(images.to(dev), ids.to(dev), mask.to(dev))["logits"]
            torch.cuda.synchronize()
            num = den = 0.0
            for k, p in single.named_parameters():
                if p.grad is not None:
                    num += (mine[k] - p.grad).double().pow(2).sum().item()
                    den += p.grad.double().pow(2).sum().item()
            res["grad_vs_single"] = (num / den) ** 0.5
            res["n_grads"] = len(mine)
            res["logits_equal_single"] = bool(torch.equal(whole, gathered))
        q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_and_inference_world2():
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r["rank"])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    print("data-parallel world 2:", res)
    assert res[0]["grad_digest"] == res[1]["grad_digest"], "ranks hold different averaged gradients"
    assert res[0]["param_digest"] == res[1]["param_digest"], "ranks diverged after the optimizer step"
    assert res[0]["logits_digest"] == res[1]["logits_digest"]
    assert res[0]["logits_equal_single"], "sharded forward differs from the single-process forward"
    assert res[0]["n_grads"] > 200 and res[0]["grad_vs_single"] <= 2e-2, res[0]["grad_vs_single"]
