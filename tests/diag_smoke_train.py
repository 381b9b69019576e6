#!/usr/bin/env python
"""Diagnostic for the smoke() training check: per-group gradient error of the linear-loss step on the sensitised
weights, for several image sizes / with and without a preceding eval forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
import synth
from oracle import train_oracle

def run(H, W, pre_eval, seed=8, lengths=(64, 40, 17, 64)):
    model = synth.build_model(0)
    sd = synth.sensitise(model.state_dict(), 1)
    model.load_state_dict(sd)
    model = model.cuda()
    B, S = len(lengths), 64
    images, ids, mask = synth.make_inputs(B, S, seed, list(lengths), H=H, W=W)
    if pre_eval:
        model.eval()
        with torch.no_grad():
            model(images[:2].cuda(), ids[:2, :32].contiguous().cuda(), mask[:2, :32].contiguous().cuda())
    for m in model.modules():
        if isinstance(m, nn.Dropout): m.p = 0.0
    mc = model.text_encoder.model_config
    mc.hidden_dropout_prob = mc.attention_probs_dropout_prob = 0.0
    model.train(); model.cnn_encoder.backbone.eval()
    R = torch.randn(B, 10, generator=torch.Generator().manual_seed(5))
    logits = model(images.cuda(), ids.cuda(), mask.cuda())["logits"]
    (logits * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    names = set(train_oracle.trainable_names(sd))
    work = {k: (v.detach().clone().float().requires_grad_(k in names) if v.is_floating_point() else v) for k, v in sd.items()}
    ref_logits = train_oracle.train_forward(work, images, ids, mask)
    (ref_logits * R).sum().backward()
    groups = {}
    for k, p in model.named_parameters():
        if k in names and work[k].grad is not None:
            g = groups.setdefault(k.split(".")[0] + ("." + k.split(".")[1] if k.startswith("cnn") else ""), [0.0, 0.0])
            g[0] += (p.grad.float().cpu() - work[k].grad).double().pow(2).sum().item()
            g[1] += work[k].grad.double().pow(2).sum().item()
    tot = sum(g[1] for g in groups.values()) ** 0.5
    print(f"H={H} W={W} pre_eval={pre_eval}: logits rel err {((logits.detach().cpu()-ref_logits.detach()).norm()/ref_logits.detach().norm()).item():.4f}",
          {k: (round((g[0] / g[1]) ** 0.5, 4), round(g[1] ** 0.5 / tot, 3)) for k, g in groups.items()},
          "global", round((sum(g[0] for g in groups.values())) ** 0.5 / tot, 4))

run(64, 64, False)
run(224, 224, False)
run(224, 224, True)
run(96, 64, False, seed=71, lengths=(64, 33, 64, 2, 17))
