"""Probe: does a weight-gradient GEMM overlap with the BatchNorm-backward passes when they run on two streams?
For one layer shape: time wgrad alone, the BN-backward trio alone, both back to back on one stream, and both
forked onto two streams, all as CUDA-graph replays (the step's execution mode).
    python tools/probe/overlap.py [batch]                                             # today's defaults: they overlap
    UB2_CO_CARVEOUT=-1 UB2_WGRAD_SMEM_KB=227 python tools/probe/overlap.py [batch]    # round-2 start: forked == serial
MEASURED (profiles/r02_side_stream.md; the runs there used a probe-time knob UB2_CARVEOUT that set one carveout for
EVERY kernel — replaced since by launch_co() on the kernels that share an SM): the pair overlaps only when both
kernels request the same shared-memory carveout, and the BatchNorm passes keep their speed only with <= 164 KB of it.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "unet-segment-pytorch_b200"))
from unet import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
SHAPES = [("inc.3", 64, 64, 512), ("down1.3", 128, 128, 256), ("down2.3", 256, 256, 128), ("down3.3", 512, 512, 64)]


def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (reps * 4)


side = torch.cuda.Stream()
print("UB2_WGRAD_SMEM_KB", os.environ.get("UB2_WGRAD_SMEM_KB", "227"), "batch", B)
for name, cin, cout, hw in SHAPES:
    a = torch.randn(B, hw, hw, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    dA = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    y = torch.randn(B, hw, hw, cout, device=dev).to(torch.bfloat16)
    ones = torch.ones(cout, device=dev)
    zeros = torch.zeros(cout, device=dev)

    def wg():
        return K.conv_wgrad(a, dy, 9)

    def bn():
        return K.bn_backward(dA, None, None, y, ones, zeros, zeros, ones, ones)

    def serial():
        wg()
        bn()

    def forked():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            p = wg()
        r = bn()
        main.wait_stream(side)
        return p, r

    tw, tb, ts, tf = graph_time(wg), graph_time(bn), graph_time(serial), graph_time(forked)
    print(f"{name:8s} wgrad {tw:6.1f} us  bn_bwd {tb:6.1f} us  serial {ts:6.1f} us  forked {tf:6.1f} us  "
          f"(ideal max {max(tw, tb):6.1f})")
