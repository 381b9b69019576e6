#!/usr/bin/env python
"""Device-timed throughput of BASELINE.json configs[1] and configs[2] (parity-test cases, not the
bench line): image_only = CNNEncoder forward, batch 256, 224x224 bf16; text_only = TextEncoder
forward, batch 256, 512 tokens with padding masks L ~ U{64..512}.

    python tools/bench_configs.py [--iters 20] [--warmup 5] > gpurun_out/configs.json
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F_IMG = 8.1769e9


def f_text(S):
    return 169.869e6 * S + 36864.0 * S * S


def timed(fn, iters, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    import torch

    import synth

    peak = 1634.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            peak = json.load(fh).get("bf16_tflops", peak)
    dev = torch.device("cuda", 0)
    model = synth.build_model(0).to(dev)
    B = args.batch
    g = torch.Generator().manual_seed(99)
    out = []
    with torch.no_grad():
        for dt in (torch.bfloat16, torch.float32):
            images = torch.randn(B, 3, 224, 224, generator=g).to(dev, dt)
            ms = timed(lambda: model.cnn_encoder(images), args.iters, args.warmup)
            v = B / (ms * 1e-3)
            out.append({"workload": f"image_only: CNNEncoder forward, batch {B}, 224x224 {str(dt)[6:]} NCHW input",
                        "value": v, "unit": "img/s", "ms": ms, "tensor_peak_frac": v * F_IMG / (peak * 1e12)})
        for S, lo in ((512, 64), (512, 512), (128, 16), (256, 32)):
            ids = torch.randint(1, 28996, (B, S), generator=g)
            lengths = torch.randint(lo, S + 1, (B,), generator=g)
            mask = (torch.arange(S).unsqueeze(0) < lengths.unsqueeze(1)).long()
            ids = ids * mask
            ids[:, 0] = 101
            ids, mask = ids.to(dev), mask.to(dev)
            ms = timed(lambda: model.text_encoder(ids, mask), args.iters, args.warmup)
            v = B / (ms * 1e-3)
            out.append({"workload": f"text_only: TextEncoder forward, batch {B}, seq {S}, masks L~U{{{lo}..{S}}}",
                        "value": v, "unit": "seq/s", "ms": ms, "live_token_fraction": float(mask.float().mean()),
                        "tensor_peak_frac": v * f_text(S) / (peak * 1e12)})
        # per-label profile of the text_only S=512 case
        ids = torch.randint(1, 28996, (B, 512), generator=g)
        lengths = torch.randint(64, 513, (B,), generator=g)
        mask = (torch.arange(512).unsqueeze(0) < lengths.unsqueeze(1)).long().to(dev)
        ids = ids.to(dev)
        eng = model.text_encoder._engine()
        eng.profile(True)
        model.text_encoder(ids, mask)
        rows = eng.profile_report()
        eng.profile(False)
        tot = sum(r["ms"] for r in rows)
        prof = [{"label": r["label"], "ms": round(r["ms"], 4), "share": round(r["ms"] / tot, 4),
                 "launches": r["launches"]} for r in sorted(rows, key=lambda r: -r["ms"])[:8]]
    print(json.dumps({"peak_tflops": peak, "configs": out, "text_only_s512_profile": prof}, indent=1))


if __name__ == "__main__":
    main()
