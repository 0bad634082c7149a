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


def _accum_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet.parallel import BatchShardedTrainer
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=1.0, bucket_mb=0.00002,
                             accumulation_steps=2)
    before = [p.detach().clone() for p in model.parameters()]
    for step in range(2):
        for micro in range(2):
            x, t = _data(2 * micro + rank)        # micro-batch index in the reference's loop order
            tr.step(x, t)
            if step == 0 and micro == 0:          # no optimizer step after the first micro-batch
                assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
    if rank == 0:
        torch.save([p.detach().clone() for p in model.parameters()], out)
    dist.destroy_process_group()


def _reference_loop(accum, steps):
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    crit = nn.CrossEntropyLoss()
    model.train()
    for _ in range(steps):
        opt.zero_grad()
        for m in range(accum):
            x, t = _data(m)
            (crit(model(x), t) / accum).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    return model


def test_two_ranks_times_two_local_micro_batches(tmp_path):
    """2 ranks x accumulation_steps 2 == the reference loop with accumulation_steps 4 (train.py:127-147).
    BatchNorm's running statistics see the micro-batches in a different order per rank, the
    parameters do not depend on them."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_accum_worker, args=(2, 29613, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    ref = _reference_loop(4, 2)
    for a, b in zip(got, ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_single_process_accumulation():
    from unet.parallel import BatchShardedTrainer
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=1.0, accumulation_steps=3)
    for _ in range(2):
        for m in range(3):
            loss = tr.step(*_data(m))
            assert loss.dim() == 0
    ref = _reference_loop(3, 2)
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-6), k   # incl. running stats


def _diverged_worker(rank, world, port, out):
    """Ranks that seeded their models differently: the trainer makes rank 0's state everybody's."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet.parallel import BatchShardedTrainer
    torch.manual_seed(100 + rank)
    model = nn.Sequential(nn.Conv2d(1, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 2, 1))
    with torch.no_grad():
        model[1].running_mean.fill_(float(rank))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=1.0)
    start = [p.detach().clone() for p in model.parameters()] + [model[1].running_mean.clone()]
    for _ in range(2):
        tr.step(*_data(rank))
    drift = model[1].running_mean.clone()       # each rank saw its own shard
    tr.sync_buffers()                           # policy "rank0"
    torch.save({"start": start, "params": [p.detach().clone() for p in model.parameters()], "drift": drift,
                "synced": model[1].running_mean.clone()}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_ranks_start_from_rank0_and_buffers_follow_the_policy(tmp_path):
    out = str(tmp_path / "r")
    mp.spawn(_diverged_worker, args=(2, 29615, out), nprocs=2, join=True)
    r0, r1 = (torch.load(f"{out}.{r}", weights_only=False) for r in range(2))
    for a, b in zip(r0["start"], r1["start"]):
        assert torch.equal(a, b)                # broadcast at construction (parameters AND buffers)
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                # identical updates afterwards
    assert not torch.equal(r0["drift"], r1["drift"])
    assert torch.equal(r0["synced"], r1["synced"]) and torch.equal(r0["synced"], r0["drift"])


def test_flush_applies_the_left_over_micro_batches():
    """An epoch of 4 micro-batches with accumulation_steps 3 (train.py:153-159): one full step, then the
    tail step on the single left-over micro-batch, un-rescaled."""
    from unet.parallel import BatchShardedTrainer
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=1.0, accumulation_steps=3)
    for m in range(4):
        tr.step(*_data(m))
    assert tr.flush() is True and tr.flush() is False
    tr.step(*_data(0))      # the next epoch starts a fresh window: gradients are zeroed first

    ref = _model()
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    crit = nn.CrossEntropyLoss()
    ref.train()
    ropt.zero_grad()
    for i in range(4):
        x, t = _data(i)
        (crit(ref(x), t) / 3).backward()
        if (i + 1) % 3 == 0:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
            ropt.step()
            ropt.zero_grad()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    ropt.step()
    ropt.zero_grad()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    x, t = _data(0)
    (crit(ref(x), t) / 3).backward()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)


# --------------------------------------------------------------------------- gradient sink + hooks, two ranks
class _SinkLinear(torch.autograd.Function):
    """What unet.ops does on the GPU, on CPU tensors: with a gradient sink installed the backward adds the
    parameter gradient straight into ``param.grad`` (the bucket view), announces it and returns None."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        ctx.param = w
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        from unet import ops
        x, w = ctx.saved_tensors
        gw = dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
        t = ops._grad_targets(ctx.param)
        if t is not None:
            t[0].add_(gw)
            ops._grads_done(ctx.param)
            gw = None
        return dy @ w, gw


class _SinkNet(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.w1 = nn.Parameter(torch.randn(16, 8) * 0.3)
        self.w2 = nn.Parameter(torch.randn(16, 16) * 0.3)
        self.w3 = nn.Parameter(torch.randn(4, 16) * 0.3)
        self.b = nn.Parameter(torch.zeros(4))          # an ordinary autograd gradient next to the sunk ones

    def forward(self, x):
        h = torch.tanh(_SinkLinear.apply(x, self.w1))
        h = torch.tanh(_SinkLinear.apply(h, self.w2))
        return _SinkLinear.apply(h, self.w3) + self.b


def _sink_data(rank):
    g = torch.Generator().manual_seed(50 + rank)
    return torch.randn(6, 8, generator=g), torch.randint(0, 4, (6,), generator=g)


def _sink_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet.parallel import BatchShardedTrainer
    model = _SinkNet()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    # tiny buckets: w3+b | w2 | w1 — a bucket announced twice would be all-reduced before backward has filled it
    tr = BatchShardedTrainer(model, nn.CrossEntropyLoss(), opt, grad_clip=0.0, bucket_mb=0.0004)
    assert len(tr.buckets) >= 2
    for _ in range(3):
        tr.step(*_sink_data(rank))
    for b in tr.buckets:
        assert b.pending == 0, b.pending          # every gradient counted exactly once
    torch.save([p.detach().clone() for p in model.parameters()], f"{out}.{rank}")
    dist.destroy_process_group()


def test_gradient_sink_with_hooks_counts_each_gradient_once(tmp_path):
    """Two ranks, kernels that write their gradients through the sink (as every unet.ops function does on the
    GPU) next to a parameter that goes through autograd: replicas stay bit-identical and equal the reference's
    accumulation loop.  Round 1 counted sunk gradients twice when world_size > 1 (sink + post-accumulate hook),
    launched the all-reduce of a bucket before backward had filled it, and the ranks drifted apart."""
    out = str(tmp_path / "r")
    mp.spawn(_sink_worker, args=(2, 29617, out), nprocs=2, join=True)
    r0, r1 = (torch.load(f"{out}.{r}", weights_only=False) for r in range(2))
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)
    ref = _SinkNet()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    crit = nn.CrossEntropyLoss()
    for _ in range(3):
        opt.zero_grad()
        for r in range(2):
            x, t = _sink_data(r)
            (crit(ref(x), t) / 2).backward()
        opt.step()
    for a, b in zip(r0, ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
