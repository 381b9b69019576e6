#!/usr/bin/env python
"""Diagnostic: gradient error of the B200 training step against the fp32 oracle, next to the error a plain
PyTorch bf16-autocast run of the SAME oracle shows against fp32 (the rounding-noise floor of bf16 training
on this case).  Prints the worst tensors.  GPU box only; not part of the test suite."""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.nn as nn
import torch.nn.functional as F

import synth
from oracle import train_oracle as T


def oracle_grads(sd, images, ids, mask, labels, autocast):
    names = set(T.trainable_names(sd))
    work = {k: (v.detach().clone().float().cuda().requires_grad_(k in names) if v.is_floating_point() else v.cuda())
            for k, v in sd.items()}
    # train_oracle builds its key bias on the CPU: move inputs, patch zeros via default device
    torch.set_default_device("cuda")
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits = T.train_forward(work, images.cuda(), ids.cuda(), mask.cuda())
        loss = F.cross_entropy(logits.float(), labels.cuda())
        loss.backward()
    finally:
        torch.set_default_device("cpu")
    return loss.item(), {k: work[k].grad.detach().float().cpu() for k in names if work[k].grad is not None}


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    case = sys.argv[1] if len(sys.argv) > 1 else "fixture"
    sd = synth.train_weights(0)
    if case == "fixture":
        B, S, lengths, seed, labels = 4, 32, [32, 20, 7, 1], 41, [3, 1, 7, 3]
    else:
        B, S, lengths, seed, labels = 16, 128, None, 43, list(range(10)) + [1, 2, 3, 4, 5, 6]
    images, ids, mask = synth.make_inputs(B, S, seed, lengths, H=64, W=64)
    labels = torch.tensor(labels)
    l32, g32 = oracle_grads(sd, images, ids, mask, labels, False)
    l16, g16 = oracle_grads(sd, images, ids, mask, labels, True)
    model = synth.build_model(0)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    mc = model.text_encoder.model_config
    mc.hidden_dropout_prob = mc.attention_probs_dropout_prob = 0.0
    model = model.cuda().train()
    model.cnn_encoder.backbone.eval()
    out = model(images.cuda(), ids.cuda(), mask.cuda())
    loss = F.cross_entropy(out["logits"], labels.cuda())
    loss.backward()
    named = dict(model.named_parameters())
    print(f"case {case}: loss fp32 {l32:.6f} autocast {l16:.6f} ours {loss.item():.6f}")
    rows = []
    tot = sum(g.double().pow(2).sum() for g in g32.values()).sqrt().item()
    num_o = num_a = 0.0
    for k, r in g32.items():
        if r.norm().item() < 1e-5 * tot:
            continue
        o = named[k].grad.float().cpu()
        a = g16[k]
        eo, ea = ((o - r).norm() / r.norm()).item(), ((a - r).norm() / r.norm()).item()
        num_o += (o - r).double().pow(2).sum().item()
        num_a += (a - r).double().pow(2).sum().item()
        rows.append((eo, ea, r.norm().item() / tot, k))
    rows.sort(reverse=True)
    print(f"global rel-L2: ours {num_o ** 0.5 / tot:.4f}  autocast {num_a ** 0.5 / tot:.4f}")
    print("ours_err  autocast_err  share_of_norm  tensor")
    for eo, ea, sh, k in rows[:25]:
        print(f"{eo:8.4f}  {ea:8.4f}  {sh:8.4f}  {k}")
    import statistics
    print("median ours", statistics.median(r[0] for r in rows), "median autocast", statistics.median(r[1] for r in rows))


if __name__ == "__main__":
    main()
