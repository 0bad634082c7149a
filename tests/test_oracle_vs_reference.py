"""CPU, build container only: the oracle against the imported, unmodified reference on fresh
random cases (skipped where /root/reference is absent, e.g. on the GPU box)."""
import os
import sys

import pytest
import torch

from oracle import unet_oracle as O

REF = os.environ.get("UNET_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "unet")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.gen_golden import import_reference
    return import_reference()


@pytest.mark.parametrize("attention,bilinear,hw", [(True, True, (33, 47)), (False, True, (40, 24)), (True, False, (32, 32))])
def test_forward_backward_random(ref, attention, bilinear, hw):
    layers, net, loss_mod, _ = ref
    cfg = dict(n_channels=3, n_classes=3, bilinear=bilinear, base_features=8, attention=attention)
    sd = O.synthetic_state_dict(9, **cfg)
    mk = dict(n_channels=3, n_classes=3, bilinear=bilinear, base_features=8)
    model = (net.AttentionUNet if attention else net.UNet)(**mk)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, *hw, generator=g)
    t = torch.randint(0, 3, (2, *hw), generator=g)
    model.train()
    out = model(x)
    loss = loss_mod.DiceBCELoss()(out, t)
    loss.backward()
    o_loss, o_logits, o_grads = O.train_grads(x, t, O.clone_state(sd), attention=attention, bilinear=bilinear)
    assert torch.allclose(o_logits, out.detach(), rtol=1e-3, atol=1e-4)
    assert abs(o_loss.item() - loss.item()) < 1e-5
    for k, p in model.named_parameters():
        assert torch.allclose(o_grads[k], p.grad, rtol=5e-3, atol=1e-5 * p.grad.abs().max().item() + 1e-9), k


def test_metrics_python_loop(ref):
    _, _, _, metrics_mod = ref
    g = torch.Generator().manual_seed(2)
    z = torch.randn(1, 4, 16, 16, generator=g)
    t = torch.randint(0, 5, (1, 16, 16), generator=g)  # label 4 is out of range -> skipped
    m = metrics_mod.SegmentationMetrics(4)
    m.update(z, t)
    assert (m.get_confusion_matrix() == O.confusion_matrix(z, t, 4)).all()
