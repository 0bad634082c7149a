"""FusedAdamW (clip_grad_norm_ + AdamW in two multi-tensor launches) against torch's own
clip_grad_norm_ + torch.optim.AdamW (scripts/train.py:141, :346-350)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 1, 3, 3), (64,), (2,), (128, 64, 3, 3), (1, 32, 1, 1), (37,), (256, 128, 3, 3), (5, 7)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in SHAPES]


def _grads(step, scale):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(s, generator=g).cuda() * scale for s in SHAPES]


@pytest.mark.parametrize("max_norm,scale", [(1.0, 3.0), (1.0, 1e-3), (0.0, 1.0)])
def test_fused_adamw_matches_torch(max_norm, scale):
    from unet.optim import FusedAdamW

    pa, pb = _params(1), _params(1)
    ref = torch.optim.AdamW(pa, lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.99), eps=1e-8)
    fused = FusedAdamW(pb, lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=max_norm,
                       write_clipped_grads=True)
    for step in range(5):
        if step == 3:   # a scheduler changes the learning rate
            for o in (ref, fused):
                o.param_groups[0]["lr"] = 3e-3
        for ps in (pa, pb):
            for p, g in zip(ps, _grads(step, scale)):
                p.grad = g.clone()
        norm = None
        if max_norm > 0:
            norm = torch.nn.utils.clip_grad_norm_(pa, max_norm)
        ref.step()
        fused.step()
        if norm is not None:
            assert torch.allclose(fused.total_grad_norm, norm, rtol=1e-5)
        for a, b in zip(pa, pb):
            assert torch.allclose(a, b, rtol=2e-5, atol=1e-6), (step, (a - b).abs().max())
            assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)
    # state_dict interchange, both directions
    sd = fused.state_dict()
    assert float(sd["state"][0]["step"]) == 5.0 and not sd["state"][0]["step"].is_cuda
    ref2 = torch.optim.AdamW(pa, lr=3e-3, weight_decay=1e-2, betas=(0.9, 0.99), eps=1e-8)
    ref2.load_state_dict(sd)
    fused2 = FusedAdamW(pb, lr=3e-3, weight_decay=1e-2, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=max_norm)
    fused2.load_state_dict(ref.state_dict())
    for ps in (pa, pb):
        for p, g in zip(ps, _grads(9, scale)):
            p.grad = g.clone()
    if max_norm > 0:
        torch.nn.utils.clip_grad_norm_(pa, max_norm)
    ref2.step()
    fused2.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-6)


def test_fused_adamw_in_cuda_graph():
    from unet.optim import FusedAdamW

    pa, pb = _params(2), _params(2)
    for ps in (pa, pb):
        for p, g in zip(ps, _grads(0, 1.0)):
            p.grad = g.clone()
    eager = FusedAdamW(pa, lr=1e-2, max_grad_norm=1.0)
    graphed = FusedAdamW(pb, lr=1e-2, max_grad_norm=1.0)
    eager.step()
    graphed.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        graphed.step()
    for lr in (1e-2, 5e-3):
        for o in (eager, graphed):
            o.param_groups[0]["lr"] = lr
        eager.step()
        graphed.sync_hyperparams()
        graph.replay()
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
