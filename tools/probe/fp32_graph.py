"""Probe: does the fp32-accuracy-mode training step capture into a CUDA graph, and what does replay gain?
MEASURED (B200, batch 4): it captures (same losses to the last bit); eager 35.35 ms/step, replayed 33.84 ms/step —
the fp32 mode is bound by its kernels (3xTF32 convolutions, three bf16 weight-gradient launches per layer), not
by the host."""
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200"))
sys.path.insert(0, ROOT)
import unet  # noqa: E402
from oracle.unet_oracle import synthetic_batch  # noqa: E402
from unet.models import AttentionUNet  # noqa: E402
from unet.optim import FusedAdamW  # noqa: E402
from unet.parallel import BatchShardedTrainer  # noqa: E402
from unet.utils.loss import DiceBCELoss  # noqa: E402

unet.set_precision("tf32")
dev = torch.device("cuda:0")
x, t = synthetic_batch(4, 512, 512, seed=1234)
x, t = x.to(dev), t.to(dev)
for graph in (False, True):
    torch.manual_seed(0)
    model = AttentionUNet(1, 2, True, 64).to(dev)
    tr = BatchShardedTrainer(model, DiceBCELoss(), FusedAdamW(model.parameters(), lr=1e-4), grad_clip=1.0,
                             cuda_graph=graph, graph_warmup=2)
    losses = [tr.step(x, t).item() for _ in range(5)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        tr.step(x, t)
    torch.cuda.synchronize()
    print("graph" if graph else "eager", f"{(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step", losses)
