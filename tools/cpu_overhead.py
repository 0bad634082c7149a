"""How much of a train step is host-side launch overhead?  Times the step on a tiny input
(GPU work negligible) next to the 512^2 batch-4 step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200")); sys.path.insert(0, ROOT)
from bench import synthetic_batch  # noqa: E402  (the bench's own generator; oracle/ is for tests only)
from unet.models import AttentionUNet
from unet.parallel import BatchShardedTrainer
from unet.utils.loss import DiceBCELoss

dev = torch.device("cuda")
torch.manual_seed(0)
model = AttentionUNet().to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=1e-4, foreach=True)
tr = BatchShardedTrainer(model, DiceBCELoss(), opt, grad_clip=1.0)
for n, hw in ((1, 64), (4, 512)):
    x, t = synthetic_batch(n, hw, hw, seed=1)
    x, t = x.to(dev), t.to(dev)
    for _ in range(5):
        tr.step(x, t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        tr.step(x, t)
    t_issue = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 20
    print(f"batch {n} {hw}x{hw}: host issue {t_issue*1e3:.2f} ms/step, wall {t_all*1e3:.2f} ms/step")
