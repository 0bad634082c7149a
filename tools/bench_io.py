"""Time the two io.cu kernels alone (CUDA events, L2 flushed between launches) and print achieved
GB/s over their algorithmic bytes: prepare 2 B read + 12 B written per pixel, predict 8 + 1."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402


def timed(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e-3


def main():
    for n in (4, 32, 128):
        px = n * 512 * 512
        img = torch.randint(0, 256, (n, 512, 512), dtype=torch.uint8, device="cuda")
        lab = torch.randint(0, 256, (n, 512, 512), dtype=torch.uint8, device="cuda")
        flg = (torch.arange(n, device="cuda") % 4).to(torch.uint8)
        x = torch.empty((n, 1, 512, 512), device="cuda")
        t = torch.empty((n, 512, 512), dtype=torch.int64, device="cuda")
        s = timed(lambda: K.prepare_batch(img, lab, flg, x=x, targets=t))
        z = torch.randn(n, 2, 512, 512, device="cuda")
        m = torch.empty((n, 512, 512), dtype=torch.uint8, device="cuda")
        p = torch.empty((n,), dtype=torch.int32, device="cuda")
        s2 = timed(lambda: K.predict_mask(z, 0.5, m, p))
        print(f"batch {n:4d}: prepare_batch {s*1e6:8.1f} us {14*px/s/1e9:7.0f} GB/s | "
              f"predict_mask {s2*1e6:8.1f} us {9*px/s2/1e9:7.0f} GB/s")


if __name__ == "__main__":
    main()
