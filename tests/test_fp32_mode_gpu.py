"""fp32 / TF32 evaluation mode (BASELINE configs[0]; north_star: logits within 1e-3 relative of the
fp32 reference).  The oracle is the fp32 restatement of the reference (pinned by tests/golden)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north_star: fp32/TF32 mode, relative error of the logits


@pytest.fixture(autouse=True)
def _tf32_mode():
    import unet
    unet.set_precision("tf32")
    yield
    unet.set_precision("bf16")


def _tf32(x):
    """round to nearest TF32 (10-bit mantissa), ties away from zero like cvt.rna"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("n,h,w,c0,c1,cout,k", [(2, 16, 16, 32, 0, 64, 3), (1, 32, 32, 64, 64, 64, 3),
                                                (2, 8, 8, 256, 0, 512, 3), (1, 64, 64, 64, 0, 32, 1),
                                                (1, 20, 24, 16, 16, 48, 3), (1, 128, 128, 64, 0, 64, 3)])
def test_conv_tf32(n, h, w, c0, c1, cout, k):
    from unet import fp32

    g = torch.Generator().manual_seed(c0 + cout)
    x0 = _tf32(torch.randn(n, h, w, c0, generator=g))
    x1 = _tf32(torch.randn(n, h, w, c1, generator=g)) if c1 else None
    wt = torch.randn(cout, c0 + c1, k, k, generator=g) / (k * (c0 + c1) ** 0.5)
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    got = fp32.conv(x0.cuda(), x1.cuda() if c1 else None, wt.cuda(), sc.cuda(), sh.cuda(), relu=True).cpu()
    x = torch.cat([x0] + ([x1] if c1 else []), dim=3).permute(0, 3, 1, 2)
    ref = F.conv2d(x.double(), _tf32(wt).double(), padding=k // 2)
    ref = torch.relu(ref * sc.view(1, -1, 1, 1).double() + sh.view(1, -1, 1, 1).double()).permute(0, 2, 3, 1).float()
    # operands are exactly representable in TF32: only fp32 accumulation order and the output rounding differ
    assert torch.allclose(got, ref, rtol=2e-3, atol=2e-4), (got - ref).abs().max()
    assert ((got - ref).norm() / ref.norm()).item() < 5e-4


@pytest.mark.parametrize("attention,base,n,hw", [(True, 32, 2, 64), (False, 32, 1, 96), (True, 64, 1, 512)])
def test_eval_forward_within_1e3(attention, base, n, hw):
    from unet.models import AttentionUNet, UNet

    cfg = dict(n_channels=1, n_classes=2, bilinear=True, base_features=base, attention=attention)
    sd = O.synthetic_state_dict(42, **cfg)
    model = (AttentionUNet if attention else UNet)(1, 2, True, base)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    x, _ = O.synthetic_batch(n, hw, hw, seed=1234)
    with torch.no_grad():
        logits = model(x.cuda()).cpu()
    ref = O.unet_forward(x, sd, attention=attention, training=False)
    err = ((logits - ref).norm() / ref.norm()).item()
    assert logits.dtype == torch.float32 and logits.shape == ref.shape
    assert err <= TOL, f"TF32 eval logits rel-L2 {err:.3e}"
    agree = (logits.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"TF32 eval {hw}x{hw} base {base}: logits rel-L2 {err:.2e}, argmax agreement {agree:.5f}")
    assert agree >= 0.999   # random-init logits are near ties on many pixels
