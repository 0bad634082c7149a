"""The other BASELINE configs (parity-test cases, not bench lines), timed for DESIGN.md:
  cfg3  plain UNet 512^2 bf16 training, batch 32 (conv-kernel isolation)
  cfg5  AttentionUNet 512^2 bf16 inference (BN folded, fused threshold + confusion counts), batch sweep
Usage: python tools/bench_configs.py [cfg3] [cfg5]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
sys.path.insert(0, ROOT)
from bench import synthetic_batch  # noqa: E402  (the bench's own generator; oracle/ is for tests only)
from unet.inference import InferenceEngine  # noqa: E402
from unet.models import AttentionUNet, UNet  # noqa: E402
from unet.optim import FusedAdamW  # noqa: E402
from unet.parallel import BatchShardedTrainer  # noqa: E402
from unet.utils.loss import DiceBCELoss  # noqa: E402
from unet.utils.metrics import SegmentationMetrics  # noqa: E402

dev = torch.device("cuda")
which = set(sys.argv[1:]) or {"cfg3", "cfg5"}


def timed(fn, iters, warm):
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


if "cfg3" in which:
    torch.manual_seed(42)
    model = UNet(1, 2, True, 64).to(dev)
    tr = BatchShardedTrainer(model, DiceBCELoss(), FusedAdamW(model.parameters(), lr=5e-5, weight_decay=1e-4),
                             grad_clip=1.0, cuda_graph=True)
    x, t = synthetic_batch(32, 512, 512, seed=1)
    x, t = x.to(dev), t.to(dev)
    ms = timed(lambda: tr.step(x, t), 10, 8)
    flops = 32 * 957_509_271_552
    print(f"cfg3 UNet train batch 32: {ms:.2f} ms/step, {32 / ms * 1e3:.1f} img/s, {flops / ms / 1e9:.0f} TFLOP/s end to end")
    del tr, model
    torch.cuda.empty_cache()

if "cfg5" in which:
    torch.manual_seed(42)
    model = AttentionUNet(1, 2, True, 64).to(dev).eval()
    metrics = SegmentationMetrics(2)
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        x, t = synthetic_batch(min(b, 32), 512, 512, seed=2)
        reps = (b + x.shape[0] - 1) // x.shape[0]
        x = x.repeat(reps, 1, 1, 1)[:b].to(dev)
        t = t.repeat(reps, 1, 1)[:b].to(dev)

        engine = InferenceEngine(model)

        def step():
            metrics.update(engine(x), t)

        ms = timed(step, 5 if b >= 64 else 20, 4)
        print(f"cfg5 AttentionUNet eval batch {b:3d}: {ms:8.2f} ms, {b / ms * 1e3:7.1f} img/s, "
              f"{b * 327_891_812_352 / ms / 1e9:6.0f} TFLOP/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
