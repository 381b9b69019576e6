import os, sys, math, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for case in ["A", "B", "C", "D", "E"]:
        r = subprocess.run([sys.executable, __file__, case], capture_output=True, text=True, timeout=120)
        print(case, "rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1], (r.stderr.strip().splitlines() or [""])[-1][:200])
    sys.exit(0)
import torch, torch.nn.functional as F
import mrd_b200
from importlib import import_module
lib = import_module("multimodal-rare-disease_b200._lib").load()
case = sys.argv[1]
BF = torch.bfloat16
st = torch.cuda.current_stream().cuda_stream
def run(N, H, W, Cin, Cout, k, out_pad):
    x = torch.randn(N, Cin, H, W, device="cuda").to(BF)
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / math.sqrt(Cin * k * k)).to(BF)
    b = torch.randn(Cout, device="cuda")
    y = torch.zeros(N, H + 2 * out_pad, W + 2 * out_pad, Cout, device="cuda", dtype=BF)
    rc = lib.mrd_conv2d_nhwc_bf16(x.permute(0, 2, 3, 1).contiguous().data_ptr(), N, H, W, Cin,
                                  w.permute(0, 2, 3, 1).contiguous().data_ptr(), Cout, k, 1, b.data_ptr(),
                                  y.data_ptr(), None, 0, out_pad, st)
    assert rc == 0, lib.mrd_last_error()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), b, padding=k // 2)
    inner = y[:, 1:-1, 1:-1] if out_pad else y
    print("max err", (inner.permute(0, 3, 1, 2).float() - ref).abs().max().item())
if case == "A": run(3, 56, 56, 64, 64, 3, 1)
if case == "B": run(3, 56, 56, 64, 64, 1, 1)
if case == "C": run(2, 28, 28, 128, 128, 1, 1)
if case == "D":
    N, H, W, Cin, Cout = 3, 56, 56, 64, 64
    x = torch.randn(N, Cin, H, W, device="cuda").to(BF)
    pad = torch.zeros(N, H + 2, W + 2, Cin, device="cuda", dtype=BF)
    pad[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(Cin * 9)).to(BF)
    b = torch.randn(Cout, device="cuda")
    y = torch.zeros(N, H, W, Cout, device="cuda", dtype=BF)
    rc = lib.mrd_conv3x3_flat_bf16(pad.data_ptr(), N, H, W, Cin, w.permute(0, 2, 3, 1).contiguous().data_ptr(), Cout, b.data_ptr(), y.data_ptr(), 0, st)
    assert rc == 0, lib.mrd_last_error()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    print("flat max err", (y.permute(0, 3, 1, 2).float() - ref).abs().max().item())
if case == "E": run(3, 56, 56, 64, 64, 3, 0)
