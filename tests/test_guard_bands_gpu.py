"""Out-of-bounds WRITES of the kernels, without compute-sanitizer: every output / workspace tensor the Python
wrappers allocate (they all go through ``torch.empty`` / ``torch.zeros`` / ``torch.empty_like``) is carved out
of a larger buffer with 4 KiB canary bands on both sides; after a full training step (forward, loss, backward)
every band must be untouched.  The caching allocator rounds requests up, so a kernel that writes a few rows
past its tensor otherwise lands in slack or in a neighbouring tensor and goes unnoticed by the parity tests.
Shapes: odd sizes (masked tiles, padded up-sampling), both precisions, the ConvTranspose2d variant."""
import contextlib

import pytest
import torch

from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
GUARD = 4096
CANARY = 0xA5


@contextlib.contextmanager
def guarded_allocations():
    bands = []
    real_empty, real_zeros, real_empty_like = torch.empty, torch.zeros, torch.empty_like

    def carve(shape, dtype, device, fill=None):
        dtype = dtype or torch.get_default_dtype()
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        buf = real_empty((GUARD + nbytes + pad + GUARD,), dtype=torch.uint8, device=device)
        buf.fill_(CANARY)
        bands.append((buf, nbytes))
        t = buf[GUARD:GUARD + nbytes].view(dtype).view(tuple(int(d) for d in shape))
        if fill is not None:
            t.fill_(fill)
        return t

    def is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def empty(*shape, dtype=None, device=None, **kw):
        if not is_cuda(device) or kw.get("pin_memory") or kw.get("memory_format") is not None:
            return real_empty(*shape, dtype=dtype, device=device, **kw)
        return carve(shape, dtype, device)

    def zeros(*shape, dtype=None, device=None, **kw):
        if not is_cuda(device):
            return real_zeros(*shape, dtype=dtype, device=device, **kw)
        return carve(shape, dtype, device, fill=0)

    def empty_like(t, **kw):
        if not t.is_cuda or kw or not t.is_contiguous():
            return real_empty_like(t, **kw)
        return carve(tuple(t.shape), t.dtype, t.device)

    torch.empty, torch.zeros, torch.empty_like = empty, zeros, empty_like
    try:
        yield bands
    finally:
        torch.empty, torch.zeros, torch.empty_like = real_empty, real_zeros, real_empty_like


def _check(bands):
    torch.cuda.synchronize()
    assert len(bands) > 50, len(bands)
    bad = 0
    for buf, nbytes in bands:
        lo, hi = buf[:GUARD], buf[GUARD + nbytes + ((-nbytes) % 256):]
        if not bool((lo == CANARY).all()) or not bool((hi == CANARY).all()):
            bad += 1
    assert bad == 0, f"{bad} of {len(bands)} guarded tensors have a damaged canary band"


@pytest.mark.parametrize("precision,bilinear,hw", [("bf16", True, (72, 88)), ("bf16", False, (64, 64)), ("tf32", True, (72, 88)),
                                                   ("bf16", True, (128, 256))])
def test_training_step_writes_stay_inside_their_tensors(precision, bilinear, hw):
    import unet
    from unet.models import AttentionUNet
    from unet.utils.loss import DeepSupervisionLoss, DiceBCELoss
    from unet.utils.metrics import SegmentationMetrics
    unet.set_precision(precision)
    try:
        torch.manual_seed(1)
        ds = precision == "bf16" and bilinear
        model = AttentionUNet(1, 2, bilinear, 32, deep_supervision=ds).cuda().train()
        x, t = O.synthetic_batch(2, hw[0], hw[1], seed=4, fg_fraction=0.05)
        x, t = x.cuda(), t.cuda()
        with guarded_allocations() as bands:
            out = model(x)
            crit = DeepSupervisionLoss(DiceBCELoss()) if ds else DiceBCELoss()
            loss = crit(out, t)
            loss.backward()
            m = SegmentationMetrics(2)
            m.update(out[0] if ds else out, t)
            _check(bands)
        assert torch.isfinite(loss).item()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    finally:
        unet.set_precision("bf16")


def test_inference_writes_stay_inside_their_tensors():
    from unet.models import AttentionUNet, UNet
    for cls in (AttentionUNet, UNet):
        torch.manual_seed(2)
        model = cls(1, 2, True, 32).cuda().eval()
        x, _ = O.synthetic_batch(3, 40, 56, seed=5)
        with guarded_allocations() as bands, torch.no_grad():
            model(x.cuda())
            _check(bands)
