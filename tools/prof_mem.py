"""One launch of each bandwidth-bound kernel at the 512^2 x 64-channel level (for ncu --set full).
Usage: python tools/prof_mem.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda"
h, c = 512, 64
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
y, dA = rnd(N, h, h, c), rnd(N, h, h, c)
low = rnd(N, h // 2, h // 2, c)
sc = torch.ones(c, device=dev); sh = torch.zeros(c, device=dev)
for _ in range(2):
    K.bn_act(y, sc, sh, True, True, want_idx=True)
    K.bn_act(y, sc, sh, True, False)
    K.bn_backward(dA, None, None, y, sc, sh, sh, sc, sc)
    K.upsample(low, h, h, h, h)
    K.upsample_bwd(y, h // 2, h // 2, h, h)
    # attention gate at the up4 level: g (N,256,256,64) -> q (N,256,256,32); x (N,512,512,64)
    ci = 32
    q, xp = rnd(N, h // 2, h // 2, ci), rnd(N, h, h, ci)
    s32 = torch.ones(ci, device=dev); z32 = torch.zeros(ci, device=dev)
    wpsi = torch.randn(ci, device=dev, generator=g)
    K.gate_upstats(q, h, h)
    psi, st = K.gate_psi(q, xp, s32, z32, s32, z32, wpsi)
    one = torch.ones(1, device=dev); zero = torch.zeros(1, device=dev)
    out, a = K.gate_apply(psi, one, zero, y)
    dx, dpsin, part = K.gate_bwd_a(dA, y, a, psi)
    coef_p = torch.ones(3, 1, device=dev)
    ds, part2 = K.gate_bwd_s(dpsin, psi, coef_p, q, xp, s32, z32, s32, z32, wpsi)
    coef = torch.ones(6, ci, device=dev)
    K.gate_bwd_xg(ds, xp, q, coef)
    # network ends
    x = torch.randn(N, 1, h, h, device=dev, generator=g)
    w0 = torch.randn(64, 1, 3, 3, device=dev, generator=g)
    K.conv_in_fwd(x, w0)
    K.conv_in_wgrad(x, dA, 64)
    wo = torch.randn(2, 64, device=dev, generator=g); bo = torch.zeros(2, device=dev)
    logits = K.outc_fwd(y, wo, bo)
    K.outc_bwd(torch.randn_like(logits), y, wo)
torch.cuda.synchronize()
print("ok")
