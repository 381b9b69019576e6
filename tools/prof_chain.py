#!/usr/bin/env python
"""The chained bottleneck tail (mrd_conv_chain_bf16) against the two launches it replaces, at layer1 / layer2 sizes
of one 512-image ResNet pass.  Prints CUDA-event times; under ncu it is the target of the DRAM-traffic capture:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
        --clock-control none -k regex:'conv_chain|conv_gemm' python tools/prof_chain.py
"""
import ctypes as C
import math
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, (N, H, W, C0, Cout, C2, pad) in {"layer1": (512, 56, 56, 64, 256, 64, 1),
                                           "layer1->2": (512, 56, 56, 64, 256, 128, 0)}.items():
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(N, H, W, C0, device=dev, generator=g).to(BF)
    ident = torch.randn(N, H, W, Cout, device=dev, generator=g).to(BF)
    w1 = (torch.randn(Cout, C0, device=dev, generator=g) / math.sqrt(C0)).to(BF)
    b1 = torch.randn(Cout, device=dev, generator=g)
    w2 = (torch.randn(C2, Cout, device=dev, generator=g) / math.sqrt(Cout)).to(BF)
    b2 = torch.randn(C2, device=dev, generator=g)
    y = torch.empty(N, H, W, Cout, device=dev, dtype=BF)
    z = torch.zeros(N, H + 2 * pad, W + 2 * pad, C2, device=dev, dtype=BF)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def chain():
        assert lib.mrd_conv_chain_bf16(x0.data_ptr(), C0, None, 0, 1, ident.data_ptr(), N, H, W, w1.data_ptr(), Cout,
                                       b1.data_ptr(), y.data_ptr(), w2.data_ptr(), C2, b2.data_ptr(), z.data_ptr(),
                                       pad, s()) == 0, lib.mrd_last_error()

    def separate():
        assert lib.mrd_conv2d_nhwc_bf16(x0.data_ptr(), N, H, W, C0, w1.data_ptr(), Cout, 1, 1, b1.data_ptr(),
                                        y.data_ptr(), ident.data_ptr(), 1, 0, s()) == 0, lib.mrd_last_error()
        assert lib.mrd_conv2d_nhwc_bf16(y.data_ptr(), N, H, W, Cout, w2.data_ptr(), C2, 1, 1, b2.data_ptr(),
                                        z.data_ptr(), None, 1, pad, s()) == 0, lib.mrd_last_error()

    gb = 2.0 * N * H * W * (C0 + 2 * Cout + C2) / 1e9
    t_sep, t_chain = timed(separate), timed(chain)
    print(f"{name}: separate {t_sep:.1f} us, chain {t_chain:.1f} us "
          f"({gb / (t_chain * 1e-6) / 1e3:.2f} TB/s on {gb:.2f} GB algorithmic)", flush=True)
