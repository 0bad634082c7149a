"""ModelEMA (one multi-tensor launch) against the reference's update arithmetic
(unet/utils/general.py:155-184), stand-alone and inside the trainer's captured step."""
import copy

import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _ref_update(ema_model, model, decay):
    """general.py:171-184, verbatim arithmetic"""
    with torch.no_grad():
        src = dict(model.named_parameters())
        for name, e in ema_model.named_parameters():
            e.data.mul_(decay).add_(src[name].data, alpha=1 - decay)
        srcb = dict(model.named_buffers())
        for name, e in ema_model.named_buffers():
            e.data.copy_(srcb[name].data)


@pytest.mark.parametrize("warmup_steps", [0, 3])
def test_model_ema_matches_reference(warmup_steps):
    from unet.models import AttentionUNet
    from unet.utils import ModelEMA

    torch.manual_seed(1)
    model = AttentionUNet(1, 2, True, 16).cuda()
    ema = ModelEMA(model, decay=0.99, warmup_steps=warmup_steps)
    ref = copy.deepcopy(model).eval()
    g = torch.Generator(device="cuda").manual_seed(2)
    for step in range(1, 6):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn(p.shape, device="cuda", generator=g) * 0.01)
            for b in model.buffers():
                if b.dtype == torch.int64:
                    b.add_(1)
                else:
                    b.add_(0.1)
        decay = min(0.99, (1 + step) / (10 + step)) if step <= warmup_steps else 0.99
        _ref_update(ref, model, decay)
        ema.update(model)
        assert ema.updates == step
    for (n, a), (_, b) in zip(ema.ema_model.state_dict().items(), ref.state_dict().items()):
        if a.dtype == torch.int64:
            assert torch.equal(a, b), n
        else:
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), n
    sd = ema.state_dict()
    assert sd["updates"] == 5 and sd["decay"] == 0.99 and set(sd) == {"ema_state_dict", "decay", "updates"}


def test_trainer_updates_ema_inside_graph():
    from unet.models import AttentionUNet
    from unet.optim import FusedAdamW
    from unet.parallel import BatchShardedTrainer
    from unet.utils import ModelEMA
    from unet.utils.loss import DiceBCELoss

    x, t = O.synthetic_batch(2, 64, 64, seed=9, fg_fraction=0.05)
    x, t = x.cuda(), t.cuda()
    results = []
    for use_graph in (False, True):
        torch.manual_seed(3)
        model = AttentionUNet(1, 2, True, 32).cuda()
        ema = ModelEMA(model, decay=0.9, warmup_steps=2)
        tr = BatchShardedTrainer(model, DiceBCELoss(), FusedAdamW(model.parameters(), lr=1e-3), grad_clip=1.0,
                                 cuda_graph=use_graph, graph_warmup=2, ema=ema)
        losses = [tr.step(x, t).item() for _ in range(6)]
        results.append((losses, [v.clone() for v in ema.ema_model.state_dict().values()], ema.updates,
                        [p.detach().clone() for p in model.parameters()],
                        [p.detach().clone() for p in ema.ema_model.parameters()]))
    (l0, e0, u0, p0, q0), (l1, e1, u1, p1, _) = results
    assert u0 == u1 == 6
    assert all(abs(a - b) <= 1e-6 * abs(a) for a, b in zip(l0, l1))
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)
    for a, b in zip(e0, e1):
        assert torch.equal(a, b)
    # the EMA really moved towards the trained weights and differs from them
    assert all(a.shape == b.shape for a, b in zip(q0, p0))
    assert sum(float((a - b).abs().sum()) for a, b in zip(q0, p0)) > 0
