#!/usr/bin/env python
"""Where does a flaky GEMM launch take its wrong values from?  Inputs are built so that every output element is an
exactly representable integer that names its own source (row in tile, column, tile number), separately for the
residual path (A = 0, out == R) and the accumulator path (R = 0, one-hot A rows, out == a weight pattern).

    python tools/stress_patterns.py [reps]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mrd_b200  # noqa: E402,F401
from importlib import import_module  # noqa: E402

lib = import_module("multimodal-rare-disease_b200._lib").load()
dev = torch.device("cuda:0")
BF = torch.bfloat16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
print("env:", {k: v for k, v in os.environ.items() if k.startswith("MRD_DEBUG")}, flush=True)


def s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch(A, W, bias, R, out, act=1):
    M, K = A.shape
    N = W.shape[0]
    out.fill_(float("nan"))
    assert lib.mrd_gemm_bf16(A.data_ptr(), K, M, K, W.data_ptr(), N, bias.data_ptr(), out.data_ptr(), N,
                             R.data_ptr() if R is not None else None, N, None, 0, act, s()) == 0, lib.mrd_last_error()
    torch.cuda.synchronize()


def report(name, out, want):
    bad = (out != want).nonzero()
    if bad.numel() == 0:
        return 0
    rows, cols = bad[:, 0], bad[:, 1]
    print(f"   {name}: {bad.shape[0]} wrong; tiles {torch.unique(rows // 128).tolist()[:8]} rows-in-tile "
          f"{torch.unique(rows % 128).tolist()[:48]} cols {torch.unique(cols).tolist()[:48]}")
    for i in range(0, min(bad.shape[0], 400), max(1, bad.shape[0] // 12)):
        r, c = rows[i].item(), cols[i].item()
        print(f"      ({r // 128}:{r % 128}, {c}) want {want[r, c].item():.0f} got {out[r, c].item():.0f}")
    return 1


def run_case(M, N, K):
    print(f"--- M={M} N={N} K={K}", flush=True)
    rr = torch.arange(M, device=dev).view(M, 1).expand(M, N)
    cc = torch.arange(N, device=dev).view(1, N).expand(M, N)
    zeroA = torch.zeros(M, K, device=dev, dtype=BF)
    W = torch.zeros(N, K, device=dev, dtype=BF)
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=BF)
    pats = {
        "residual = row in tile": (rr % 128).to(BF),
        "residual = column % 256": (cc % 256).to(BF),
        "residual = tile % 256": ((rr // 128) % 256).to(BF),
        "residual = CTA round (tile // 148)": ((rr // 128) // 148).to(BF),
    }
    nbad = 0
    for name, R in pats.items():
        R = R.contiguous()
        for r in range(reps):
            launch(zeroA, W, bias, R, out)
            nbad += report(f"{name} rep {r}", out, R)
    # accumulator path: A rows one-hot at k = row % K; W[c, k] is the pattern -> out[r, c] = W[c, r % K]
    A = torch.zeros(M, K, device=dev, dtype=BF)
    A[torch.arange(M, device=dev), torch.arange(M, device=dev) % K] = 1
    zeroR = torch.zeros(M, N, device=dev, dtype=BF)
    kk = torch.arange(K, device=dev).view(1, K).expand(N, K)
    cn = torch.arange(N, device=dev).view(N, 1).expand(N, K)
    for name, Wp in {"acc = column % 256": (cn % 256).to(BF), "acc = row % K": (kk % 256).to(BF)}.items():
        Wp = Wp.contiguous()
        want = Wp.float().t()[torch.arange(M, device=dev) % K].to(BF)
        for with_res in (True, False):
            for r in range(reps):
                launch(A, Wp, bias, zeroR if with_res else None, out)
                nbad += report(f"{name} (residual {'zero' if with_res else 'none'}) rep {r}", out, want)
    print(f"   {nbad} bad launches", flush=True)
    return nbad


total = run_case(200704, 256, 64) + run_case(100352, 512, 128)
sys.exit(1 if total else 0)
