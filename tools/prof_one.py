"""Run a few launches of one conv shape (for ncu --set full captures).
Usage: python tools/prof_one.py <fwd|wgrad> N H C0 C1 Cout taps"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

kind = sys.argv[1]
n, h, c0, c1, cout, taps = map(int, sys.argv[2:8])
dev = "cuda"
x0 = torch.randn(n, h, h, c0, device=dev).bfloat16()
x1 = torch.randn(n, h, h, c1, device=dev).bfloat16() if c1 else None
dy = torch.randn(n, h, h, cout, device=dev).bfloat16()
wf = torch.randn(cout, taps, c0 + c1, device=dev).bfloat16()
for _ in range(3):
    if kind == "fwd":
        K.conv_fwd(x0, wf, taps, x1=x1, stats=True)
    else:
        K.conv_wgrad(x0, dy, taps, x1=x1)
torch.cuda.synchronize()
print("ok")
