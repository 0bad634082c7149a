"""Per-pass timing of the bandwidth-bound kernels at the network's real level shapes (CUDA events, hot loop;
compare builds with UB2_LIB=<other .so>).  Usage: python tools/bench_mem.py [batch] [gate|bn|up|heads ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

args = [a for a in sys.argv[1:]]
N = int(args[0]) if args and args[0].isdigit() else 4
which = set(a for a in args if not a.isdigit()) or {"gate", "bn", "up", "heads"}
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
nb = K._nbytes
TOT = {}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def line(group, name, us, nbytes):
    TOT[group] = TOT.get(group, 0.0) + us
    print(f"  {name:34s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s  ({nbytes / 1e6:7.1f} MB)")


if "gate" in which:
    print(f"attention gate passes, batch {N}")
    for lvl, h, ci, cx in (("up1", 64, 256, 512), ("up2", 128, 128, 256), ("up3", 256, 64, 128), ("up4", 512, 32, 64)):
        hin = h // 2
        q, xp, x = bf(N, hin, hin, ci), bf(N, h, h, ci), bf(N, h, h, cx)
        dout = bf(N, h, h, cx)
        vec = lambda: torch.rand(ci, device=dev) + 0.5
        sg, hg, sx, hx, wpsi = vec(), vec() - 1, vec(), vec() - 1, vec() - 1
        one = torch.ones(1, device=dev)
        line("gate", f"{lvl} upstats", timeit(lambda: K.gate_upstats(q, h, h)), nb(q))
        psi, _ = K.gate_psi(q, xp, sg, hg, sx, hx, wpsi)
        line("gate", f"{lvl} psi", timeit(lambda: K.gate_psi(q, xp, sg, hg, sx, hx, wpsi)), nb(q, xp, psi))
        out, a = K.gate_apply(psi, one, one, x)
        line("gate", f"{lvl} apply", timeit(lambda: K.gate_apply(psi, one, one, x)), nb(psi, x, out, a))
        dx, dpsin, _ = K.gate_bwd_a(dout, x, a, psi)
        line("gate", f"{lvl} bwd_a", timeit(lambda: K.gate_bwd_a(dout, x, a, psi)), nb(dout, x, a, psi, dx, dpsin))
        cpsi = torch.tensor([1.0, 0.1, 0.01], device=dev)
        ds, _ = K.gate_bwd_s(dpsin, psi, cpsi, q, xp, sg, hg, sx, hx, wpsi)
        line("gate", f"{lvl} bwd_s", timeit(lambda: K.gate_bwd_s(dpsin, psi, cpsi, q, xp, sg, hg, sx, hx, wpsi)),
             nb(dpsin, psi, q, xp, ds))
        coef = torch.rand(6, ci, device=dev)
        dxp, dgup = K.gate_bwd_xg(ds, xp, q, coef)
        line("gate", f"{lvl} bwd_xg", timeit(lambda: K.gate_bwd_xg(ds, xp, q, coef)), nb(ds, xp, q, dxp, dgup))
        line("gate", f"{lvl} upsample_bwd(dgup)", timeit(lambda: K.upsample_bwd(dgup, hin, hin, h, h)), nb(dgup) * 1.25)
        del q, xp, x, dout, psi, out, a, dx, dpsin, ds, dxp, dgup

if "bn" in which:
    print(f"BatchNorm passes, batch {N}")
    for lvl, h, c, pool in (("inc/up4 512x64", 512, 64, True), ("down1/up3 256x128", 256, 128, True),
                            ("down2/up2 128x256", 128, 256, True), ("down3 64x512", 64, 512, True),
                            ("down4 32x512", 32, 512, False), ("up1.3 64x256", 64, 256, False)):
        y, dA = bf(N, h, h, c), bf(N, h, h, c)
        sc, sh = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        line("bn", f"{lvl} bn_act", timeit(lambda: K.bn_act(y, sc, sh, True, False)), 2 * nb(y))
        line("bn", f"{lvl} bn_backward (3 launches)", timeit(lambda: K.bn_backward(dA, None, None, y, sc, sh, sh, sc, sc)), 5 * nb(y))
        if pool:
            dP = bf(N, h // 2, h // 2, c)
            _, p, pidx = K.bn_act(y, sc, sh, True, True, want_idx=True)
            line("bn", f"{lvl} bn_act+pool", timeit(lambda: K.bn_act(y, sc, sh, True, True, want_idx=True)), 2 * nb(y) + nb(p, pidx))
            line("bn", f"{lvl} bn_backward+pool", timeit(lambda: K.bn_backward(dA, dP, pidx, y, sc, sh, sh, sc, sc)),
                 5 * nb(y) + 2 * nb(dP, pidx))
        del y, dA

if "up" in which:
    print(f"bilinear 2x up-sampling, batch {N}")
    for lvl, h, c in (("up1 32->64 x512", 32, 512), ("up2 64->128 x256", 64, 256), ("up3 128->256 x128", 128, 128),
                      ("up4 256->512 x64", 256, 64)):
        x = bf(N, h, h, c)
        up = K.upsample(x, 2 * h, 2 * h, 2 * h, 2 * h)
        line("up", f"{lvl} fwd", timeit(lambda: K.upsample(x, 2 * h, 2 * h, 2 * h, 2 * h)), nb(x, up))
        line("up", f"{lvl} bwd", timeit(lambda: K.upsample_bwd(up, h, h, 2 * h, 2 * h)), nb(x, up))

if "heads" in which:
    print(f"stem / output head / loss, batch {N}")
    x = torch.randn(N, 1, 512, 512, device=dev)
    w = torch.randn(64, 1, 3, 3, device=dev)
    y, _ = K.conv_in_fwd(x, w)
    line("heads", "conv_in_fwd", timeit(lambda: K.conv_in_fwd(x, w)), nb(x, y))
    dy = bf(N, 512, 512, 64)
    line("heads", "conv_in_wgrad", timeit(lambda: K.conv_in_wgrad(x, dy, 64)), nb(x, dy))
    a = bf(N, 512, 512, 64)
    wo, bo = torch.randn(2, 64, device=dev), torch.randn(2, device=dev)
    logits = K.outc_fwd(a, wo, bo)
    line("heads", "outc_fwd", timeit(lambda: K.outc_fwd(a, wo, bo)), nb(a, logits))
    dl = torch.randn_like(logits)
    line("heads", "outc_bwd", timeit(lambda: K.outc_bwd(dl, a, wo)), nb(dl, a, a))
    t = torch.randint(0, 2, (N, 512, 512), device=dev)
    line("heads", "seg_stats", timeit(lambda: K.seg_stats(logits, t)), nb(logits, t))
    coef = torch.rand(N, 3, 2, device=dev)
    line("heads", "seg_stats_bwd", timeit(lambda: K.seg_stats_bwd(logits, t, coef)), nb(logits, t, logits))

print("totals (us):", {k: round(v, 1) for k, v in TOT.items()})
