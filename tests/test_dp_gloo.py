"""CPU, world_size 2 over gloo: the batch-sharded trainer reproduces the reference's
gradient-accumulation arithmetic (scripts/train.py:127-147): two ranks with one micro-batch each
== one process accumulating the same two micro-batches with loss / 2."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(1, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 2, 1))


def _data(rank):
    g = torch.Generator().manual_seed(10 + rank)
    return torch.randn(2, 1, 8, 8, generator=g), torch.randint(0, 2, (2, 8, 8), generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet.parallel import BatchShardedTrainer
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=1.0, bucket_mb=0.00002)
    assert len(tr.buckets) > 1
    x, t = _data(rank)
    for _ in range(2):
        loss = tr.step(x, t)
    if rank == 0:
        torch.save({"params": [p.detach().clone() for p in model.parameters()], "loss": loss}, out)
    dist.destroy_process_group()


def test_two_ranks_equal_accumulation(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    # reference loop: accumulation_steps = 2, same data, same order of optimizer steps
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    crit = nn.CrossEntropyLoss()
    model.train()
    for _ in range(2):
        opt.zero_grad()
        for r in range(2):
            x, t = _data(r)
            (crit(model(x), t) / 2).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    for a, b in zip(got["params"], model.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
