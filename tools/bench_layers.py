"""Per-layer timing of the conv kernels and the bandwidth-bound passes (CUDA events).
Usage: python tools/bench_layers.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
LAYERS = [  # name, H, C0, C1, Cout, taps
    ("inc.3", 512, 64, 0, 64, 9), ("down1.0", 256, 64, 0, 128, 9), ("down1.3", 256, 128, 0, 128, 9),
    ("down2.0", 128, 128, 0, 256, 9), ("down2.3", 128, 256, 0, 256, 9), ("down3.0", 64, 256, 0, 512, 9),
    ("down3.3", 64, 512, 0, 512, 9), ("down4.0", 32, 512, 0, 512, 9), ("up1.0", 64, 512, 512, 512, 9),
    ("up1.3", 64, 512, 0, 256, 9), ("up2.0", 128, 256, 256, 256, 9), ("up2.3", 128, 256, 0, 128, 9),
    ("up3.0", 256, 128, 128, 128, 9), ("up3.3", 256, 128, 0, 64, 9), ("up4.0", 512, 64, 64, 64, 9),
    ("up4.3", 512, 64, 0, 64, 9), ("gate1.Wx", 64, 512, 0, 256, 1), ("gate4.Wx", 512, 64, 0, 32, 1),
]


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
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
    dev = "cuda"
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = 0.0
    print(f"batch {N}")
    print(f"{'layer':10s} {'GFLOP':>8s} | {'fwd ms':>8s} {'TF/s':>6s} | {'dgrad ms':>8s} {'TF/s':>6s} | {'wgrad ms':>8s} {'TF/s':>6s}")
    for name, h, c0, c1, cout, taps in LAYERS:
        cin = c0 + c1
        x0 = torch.randn(N, h, h, c0, device=dev).bfloat16()
        x1 = torch.randn(N, h, h, c1, device=dev).bfloat16() if c1 else None
        dy = torch.randn(N, h, h, cout, device=dev).bfloat16()
        wf = torch.randn(cout, taps, cin, device=dev).bfloat16()
        wd = torch.randn(cin, taps, cout, device=dev).bfloat16()
        d0 = torch.empty(N, h, h, c0, device=dev, dtype=torch.bfloat16)
        d1 = torch.empty(N, h, h, c1, device=dev, dtype=torch.bfloat16) if c1 else None
        fl = 2.0 * N * h * h * cout * taps * cin
        t_f = timeit(lambda: K.conv_fwd(x0, wf, taps, x1=x1, stats=True))
        t_d = timeit(lambda: K.conv_fwd(dy, wd, taps, out=d0, out1=d1, split=c0))
        t_w = timeit(lambda: K.conv_wgrad(x0, dy, taps, x1=x1))
        tot["fwd"] += t_f; tot["dgrad"] += t_d; tot["wgrad"] += t_w; totf += fl
        print(f"{name:10s} {fl / 1e9:8.1f} | {t_f:8.3f} {fl / t_f / 1e9:6.0f} | {t_d:8.3f} {fl / t_d / 1e9:6.0f} | "
              f"{t_w:8.3f} {fl / t_w / 1e9:6.0f}")
    for k, v in tot.items():
        print(f"total {k}: {v:.3f} ms  {totf / v / 1e9:.0f} TFLOP/s")
    # bandwidth-bound passes at the 512^2 x 64 level
    h, c = 512, 64
    y = torch.randn(N, h, h, c, device=dev).bfloat16()
    dA = torch.randn(N, h, h, c, device=dev).bfloat16()
    dP = torch.randn(N, h // 2, h // 2, c, device=dev).bfloat16()
    sc = torch.ones(c, device=dev); sh = torch.zeros(c, device=dev)
    nbytes = y.numel() * 2
    t = timeit(lambda: K.bn_act(y, sc, sh, True, True))
    print(f"bn_act+pool   {t:.3f} ms  {2.25 * nbytes / t / 1e6:.0f} GB/s")
    t = timeit(lambda: K.bn_backward(dA, None, None, y, sc, sh, sh, sc, sc))
    print(f"bn_backward   {t:.3f} ms  {5 * nbytes / t / 1e6:.0f} GB/s (reduce+finalize+apply)")
    pidx = K.bn_act(y, sc, sh, True, True, want_idx=True)[2]
    t = timeit(lambda: K.bn_backward(dA, dP, pidx, y, sc, sh, sh, sc, sc))
    print(f"bn_backward+pool {t:.3f} ms  {5.5 * nbytes / t / 1e6:.0f} GB/s")
    t = timeit(lambda: K.upsample(dP, h, h, h, h))
    print(f"upsample fwd  {t:.3f} ms  {1.25 * nbytes / t / 1e6:.0f} GB/s")
    t = timeit(lambda: K.upsample_bwd(y, h // 2, h // 2, h, h))
    print(f"upsample bwd  {t:.3f} ms  {1.25 * nbytes / t / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
