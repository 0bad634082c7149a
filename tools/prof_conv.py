"""A few launches of representative conv shapes (for ncu --set full): forward with BN statistics,
data gradient and weight gradient.  Usage: python tools/prof_conv.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
SHAPES = [("inc.3", 512, 64, 0, 64), ("up4.0", 512, 64, 64, 64), ("down1.3", 256, 128, 0, 128),
          ("down2.3", 128, 256, 0, 256), ("up1.0", 64, 512, 512, 512), ("down4.0", 32, 512, 0, 512)]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
for rep in range(2):
    for name, h, c0, c1, cout in SHAPES:
        x0 = rnd(N, h, h, c0)
        x1 = rnd(N, h, h, c1) if c1 else None
        dy = rnd(N, h, h, cout)
        wf = rnd(cout, 9, c0 + c1)
        wd = rnd(c0 + c1, 9, cout)
        d0 = torch.empty(N, h, h, c0, device=dev, dtype=torch.bfloat16)
        d1 = torch.empty(N, h, h, c1, device=dev, dtype=torch.bfloat16) if c1 else None
        K.conv_fwd(x0, wf, 9, x1=x1, stats=True)
        K.conv_fwd(dy, wd, 9, out=d0, out1=d1, split=c0)
        K.conv_wgrad(x0, dy, 9, x1=x1)
torch.cuda.synchronize()
print("ok")
