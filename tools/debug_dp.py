"""torchrun -N2 debug: after ONE eager trainer step, are the all-reduced gradient buckets identical on the ranks?"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "unet-segment-pytorch_b200")); sys.path.insert(0, ROOT)
from bench import synthetic_batch
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from unet.models import AttentionUNet
from unet.optim import FusedAdamW
from unet.parallel import BatchShardedTrainer
from unet.utils.loss import DiceBCELoss
torch.manual_seed(42)
model = AttentionUNet(1, 2, True, 32).to(dev)
opt = FusedAdamW(model.parameters(), lr=0.0, weight_decay=0.0)   # lr 0: parameters stay, gradients remain in the buckets
tr = BatchShardedTrainer(model, DiceBCELoss(), opt, grad_clip=0.0)
events = []
orig = tr._bucket_ready
def traced(bucket, n=1, param=None):
    before = bucket.pending
    orig(bucket, n, param)
    events.append((tr.buckets.index(bucket), before, bucket.pending, len(bucket.deferred), bucket.work is not None))
tr._bucket_ready = traced
x, t = synthetic_batch(2, 128, 128, seed=1234 + rank)
tr.step(x.to(dev), t.to(dev))
torch.cuda.synchronize()
bad = []
for bi, b in enumerate(tr.buckets):
    ref = b.flat.clone(); dist.broadcast(ref, src=0)
    d = (b.flat - ref).abs()
    if d.max().item() != 0:
        # which params
        for p in b.params:
            g = p.grad; r = g.clone(); dist.broadcast(r, src=0)
            if (g - r).abs().max().item() != 0:
                name = [n for n, q in model.named_parameters() if q is p][0]
                bad.append((bi, name, (g - r).abs().max().item(), g.abs().max().item()))
    else:
        for p in b.params:
            r = p.grad.clone(); dist.broadcast(r, src=0)
if rank == 0:
    print("buckets", [(len(b.params), b.flat.numel()) for b in tr.buckets])
    print("first all-reduce launches (bucket, pending before, after, deferred, launched):", [e for e in events if e[4]][:6])
    print("n events", len(events), "params", sum(len(b.params) for b in tr.buckets))
print(f"rank {rank}: differing grads {len(bad)}", bad[:12])
dist.destroy_process_group()
