"""GPU parity of the fused loss statistics and the confusion-matrix kernel vs the oracle.

Loss: fp32 arithmetic on both sides, different summation order only: value rel 1e-5,
gradient rel-L2 1e-4.  Metrics: integer counts — bit-exact."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def _logits_targets(n, c, h, w, seed, fg=0.05):
    g = torch.Generator().manual_seed(seed)
    z = 2.0 * torch.randn(n, c, h, w, generator=g)
    t = (torch.rand(n, h, w, generator=g) < fg).long()
    if c > 2:
        t = t * torch.randint(1, c, (n, h, w), generator=g)
    return z, t


@pytest.mark.parametrize("n,c,h,w", [(4, 2, 64, 64), (2, 2, 37, 53), (3, 4, 32, 32), (1, 2, 512, 512)])
def test_dice_bce_matches_oracle(n, c, h, w):
    from unet.utils.loss import DiceBCELoss
    z, t = _logits_targets(n, c, h, w, 1)
    if n > 1:
        t[0] = 0  # an image without foreground exercises the +1e-6 branch (loss.py:139)
    zr = z.clone().requires_grad_(True)
    ref = O.dice_bce_loss(zr, t)
    ref.backward()
    zc = z.cuda().requires_grad_(True)
    got = DiceBCELoss()(zc, t.cuda())
    got.backward()
    assert got.dim() == 0
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-6
    err = (zc.grad.cpu() - zr.grad).norm() / zr.grad.norm()
    assert err.item() <= 1e-4, f"dlogits rel-L2 {err.item():.3e}"


def test_loss_variants_and_scaling():
    from unet.utils.loss import BalancedCELoss, DeepSupervisionLoss, DiceBCELoss, DiceLoss
    z, t = _logits_targets(2, 2, 48, 48, 2)
    zc, tc = z.cuda(), t.cuda()
    assert abs(DiceLoss()(zc, tc).item() - O.dice_loss(z, t).item()) < 1e-5
    assert abs(BalancedCELoss(0.3)(zc, tc).item() - O.balanced_ce(z, t, 0.3).item()) < 1e-5
    none = DiceLoss(reduction='none', ignore_background=False)(zc, tc)
    assert none.shape == (2, 2)
    ds = DeepSupervisionLoss(DiceBCELoss())([zc, zc * 0.5, zc * 0.25, zc * 0.1], tc)
    ref = O.deep_supervision_loss([z, z * 0.5, z * 0.25, z * 0.1], t)
    assert abs(ds.item() - ref.item()) < 1e-4
    # loss / accumulation_steps (train.py:133): the upstream scale reaches the pixel gradients
    a = zc.clone().requires_grad_(True)
    (DiceBCELoss()(a, tc) / 8).backward()
    b = z.clone().requires_grad_(True)
    (O.dice_bce_loss(b, t) / 8).backward()
    assert ((a.grad.cpu() - b.grad).norm() / b.grad.norm()).item() < 1e-4


@pytest.mark.parametrize("c,ignore", [(2, None), (3, None), (3, 2), (5, 255)])
def test_confusion_matrix_bit_exact(c, ignore):
    from unet.utils.metrics import SegmentationMetrics
    g = torch.Generator().manual_seed(3)
    n, h, w = 3, 61, 47
    z = torch.randn(n, c, h, w, generator=g)
    z[:, :, :4] = 0.25  # exact ties -> first class wins
    t = torch.randint(0, c, (n, h, w), generator=g)
    if ignore is not None:
        t[torch.rand(n, h, w, generator=g) < 0.1] = ignore
    m = SegmentationMetrics(num_classes=c, ignore_index=ignore)
    m.update(z.cuda(), t.cuda())
    m.update(z.argmax(1).cuda(), t.cuda())  # class-index input path
    ref = 2 * O.confusion_matrix(z, t, c, ignore)
    assert m.confusion_matrix.dtype == np.int64
    assert np.array_equal(m.confusion_matrix, ref)
    got, want = m.compute(), O.metrics_from_confusion(ref)
    for k in ("pixel_accuracy", "mean_iou", "mean_dice"):
        assert got[k] == want[k]
    m.reset()
    assert m.compute()["mean_iou"] == 0.0


def test_iou_dice_helpers_and_threshold():
    from unet.utils.metrics import SegmentationMetrics, compute_dice, compute_iou
    z, t = _logits_targets(2, 2, 40, 40, 4, fg=0.3)
    assert torch.allclose(compute_iou(z.cuda(), t.cuda()).cpu(), O.iou_per_class(z, t), atol=1e-6)
    assert torch.allclose(compute_dice(z.cuda(), t.cuda()).cpu(), O.dice_per_class(z, t), atol=1e-6)
    # fused softmax[:,1] > thr (scripts/predict.py:155-159)
    m = SegmentationMetrics(2)
    m.update(z.cuda(), t.cuda(), threshold=0.7)
    pred = (torch.softmax(z, 1)[:, 1] > 0.7).long()
    assert np.array_equal(m.confusion_matrix, O.confusion_matrix(pred, t, 2))


def test_large_confusion_property():
    """Full-size (32 x 512^2) check through a size-independent property: the histogram sums to
    the pixel count and its marginals equal bincounts computed independently."""
    from unet.utils.metrics import SegmentationMetrics
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.randint(0, 2, (32, 512, 512), generator=g, device="cuda")
    t = torch.randint(0, 2, (32, 512, 512), generator=g, device="cuda")
    m = SegmentationMetrics(2)
    m.update(p, t)
    cm = m.confusion_matrix
    assert cm.sum() == 32 * 512 * 512
    assert np.array_equal(cm.sum(0), torch.bincount(p.flatten(), minlength=2).cpu().numpy())
    assert np.array_equal(cm.sum(1), torch.bincount(t.flatten(), minlength=2).cpu().numpy())
