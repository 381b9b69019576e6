#!/usr/bin/env python
"""One CNNEncoder forward of 512 images (one ResNet pass at the bench's micro-batch size) - the target of the
ncu captures of the stem / layer1 / layer2 kernels:

    ncu --set full --import-source on --clock-control none -k regex:'conv_gemm|maxpool|repack' -c 16 \
        -o gpurun_out/prof_cnn python tools/prof_cnn_pass.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mrd_b200  # noqa: E402,F401
from importlib import import_module  # noqa: E402

synth = import_module("multimodal-rare-disease_b200.synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
model = synth.build_model(0).to("cuda:0")
x = torch.randn(B, 3, 224, 224, device="cuda:0")
with torch.no_grad():
    y = model.cnn_encoder(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().mean()))
