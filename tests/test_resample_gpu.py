"""Bilinear (align_corners=True) up-sampling kernels against ATen's own forward and autograd
(layers.py:78, :98-102, :183)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# n, hin, win, hu, wu, Ho, Wo, C
CASES = [
    (2, 32, 32, 64, 64, 64, 64, 64),      # exact 2x, tiled transpose
    (1, 64, 128, 128, 256, 128, 256, 32),
    (2, 37, 50, 74, 100, 75, 101, 16),    # F.pad to an odd skip size
    (1, 8, 8, 16, 16, 16, 16, 48),        # tiny: staged region would not fit -> direct kernel
    (2, 16, 16, 33, 33, 33, 33, 24),      # F.interpolate to an arbitrary size, C not a multiple of 16
    (1, 20, 24, 47, 55, 47, 55, 64),      # gate-style resize (scale ~0.41)
    (1, 1, 1, 2, 2, 2, 2, 16),
    (1, 40, 36, 90, 80, 90, 80, 32),      # scale ~0.44, tall enough for the column-strip kernels
    (3, 48, 32, 96, 64, 97, 66, 64),      # strips with a remainder (48 = 3 x 16), F.pad border
]


def _ref_up(x, hu, wu, ho, wo):
    y = F.interpolate(x, size=(hu, wu), mode="bilinear", align_corners=True)
    dy, dx = ho - hu, wo - wu
    return F.pad(y, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])


@pytest.mark.parametrize("n,hin,win,hu,wu,ho,wo,c", CASES)
def test_upsample_fwd_bwd(n, hin, win, hu, wu, ho, wo, c):
    from unet import kernels as K

    g = torch.Generator().manual_seed(hin * 131 + c)
    x = torch.randn(n, hin, win, c, generator=g).bfloat16()
    dout = torch.randn(n, ho, wo, c, generator=g).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = _ref_up(xr, hu, wu, ho, wo)
    ref.backward(dout.float().permute(0, 3, 1, 2))
    got = K.upsample(x.cuda(), hu, wu, ho, wo).float().cpu()
    assert torch.allclose(got, ref.detach().permute(0, 2, 3, 1), rtol=2 ** -7, atol=2 ** -7)
    ref_g = xr.grad.permute(0, 2, 3, 1)
    got_g = K.upsample_bwd(dout.cuda(), hin, win, hu, wu).float().cpu()
    assert torch.allclose(got_g, ref_g, rtol=2 ** -7, atol=2 ** -6), (got_g - ref_g).abs().max()
    # accumulate into an existing gradient
    base = torch.randn(n, hin, win, c, generator=g).bfloat16()
    acc = base.clone().cuda()
    K.upsample_bwd(dout.cuda(), hin, win, hu, wu, into=acc)
    assert torch.allclose(acc.float().cpu(), ref_g + base.float(), rtol=2 ** -6, atol=2 ** -5)


@pytest.mark.parametrize("n,c,h,w,ho,wo", [
    (2, 2, 64, 64, 512, 512),     # ds_out3: x8
    (2, 2, 128, 128, 512, 512),   # ds_out2: x4
    (1, 2, 256, 256, 512, 512),   # ds_out1: x2
    (3, 3, 7, 9, 20, 31),         # odd sizes, any ratio
    (1, 2, 1, 5, 4, 5),           # a single source row; unchanged width
    (2, 1, 6, 6, 1, 1),           # to a single pixel (scale 0)
    (1, 2, 16, 16, 8, 12),        # down-sampling
])
def test_resize_logits_fwd_bwd(n, c, h, w, ho, wo):
    """ops.resize_logits == F.interpolate(bilinear, align_corners=True) on fp32 NCHW, forward and
    backward (ATen's CUDA kernel as the fp32 reference; gradient in gather form, deterministic)."""
    from unet import ops
    g = torch.Generator().manual_seed(h * 100 + wo)
    x = torch.randn(n, c, h, w, generator=g).cuda()
    dy = torch.randn(n, c, ho, wo, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=True)
    ref.backward(dy)
    xo = x.clone().requires_grad_(True)
    out = ops.resize_logits(xo, (ho, wo))
    out.backward(dy)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    scale = max(1.0, (ho / h) * (wo / w))
    assert torch.allclose(xo.grad, xr.grad, rtol=1e-4, atol=1e-5 * scale)
    out2 = ops.resize_logits(x, (ho, wo))
    assert torch.equal(out2, out.detach())


@pytest.mark.parametrize("n,cin,h,w,ho,wo", [(2, 128, 8, 8, 16, 16), (1, 64, 6, 9, 13, 19), (2, 256, 16, 16, 32, 32)])
def test_conv_transpose2x2_and_pad_vs_torch(n, cin, h, w, ho, wo):
    """nn.ConvTranspose2d(C, C/2, 2, 2) + F.pad (layers.py:81, :98-102) as a tensor-core 1x1 convolution +
    the pixel-shuffle kernels of csrc/shuffle.cu: output, input gradient, weight gradient and bias gradient
    against torch's fp32 transposed convolution of the same bf16-rounded operands."""
    import torch.nn.functional as F
    from unet import ops

    cout = cin // 2
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cin, cout, 2, 2, generator=g) / cin ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    dy = torch.randn(n, cout, ho, wo, generator=g).bfloat16().float()

    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.conv_transpose2d(xr, wr, br, stride=2)
    py, px = ho - 2 * h, wo - 2 * w
    y = F.pad(y, [px // 2, px - px // 2, py // 2, py - py // 2])
    y.backward(dy)

    xc = x.cuda().requires_grad_(True)
    wc = wt.cuda().requires_grad_(True)
    bc = b.cuda().requires_grad_(True)
    out = ops.conv_transpose2x2(xc, wc, bc, ho, wo)
    assert out.shape == (n, cout, ho, wo)
    out.backward(dy.cuda().to(out.dtype))

    def rel(a, r):
        return ((a.float().cpu() - r).norm() / (r.norm() + 1e-12)).item()

    assert rel(out, y.detach()) <= 4e-3            # bf16 storage of the result
    assert rel(xc.grad, xr.grad) <= 4e-3
    assert rel(wc.grad, wr.grad) <= 1e-3           # fp32 accumulation of bf16 products
    assert rel(bc.grad, br.grad) <= 1e-4           # sums of bf16 values in fp32 / fp64
    # the padding ring is exactly zero
    if py or px:
        mask = torch.ones(ho, wo, dtype=torch.bool)
        mask[py // 2: py // 2 + 2 * h, px // 2: px // 2 + 2 * w] = False
        assert (out.float().cpu()[:, :, mask] == 0).all()


# n, hin, win, H, W, C — the statistics of up(q) computed on the low-resolution tensor (gate.cu,
# gate_upstats_lowres_kernel) against sums over ATen's own interpolation (layers.py:98-102 feeding :107's BN)
UPSTATS = [
    (4, 32, 32, 64, 64, 256),      # up1 level, two channel groups per thread
    (2, 64, 64, 128, 128, 128),
    (1, 256, 256, 512, 512, 32),
    (2, 37, 50, 75, 101, 16),      # ragged target size
    (1, 20, 24, 47, 55, 64),
    (1, 1, 1, 2, 2, 16),           # degenerate: every pixel is the one source pixel
    (1, 1, 7, 3, 13, 16),
    (3, 5, 1, 9, 4, 48),
]


@pytest.mark.parametrize("n,hin,win,h,w,c", UPSTATS)
def test_gate_upstats_closed_form(n, hin, win, h, w, c):
    from unet import kernels as K
    torch.manual_seed(hin * 131 + w)
    dev = torch.device("cuda:0")
    q = (torch.randn(n, c, hin, win, device=dev) * 1.5 + 0.3).to(torch.bfloat16)
    q_nhwc = K.empty_nhwc(n, hin, win, c, dev)
    q_nhwc.copy_(q.permute(0, 2, 3, 1))
    partials = K.gate_upstats(q_nhwc, h, w)
    got = partials.sum(0)                                          # [2][C], fp64
    up = F.interpolate(q.double(), size=(h, w), mode="bilinear", align_corners=True)
    want = torch.stack([up.sum((0, 2, 3)), (up * up).sum((0, 2, 3))])
    scale = torch.stack([up.abs().sum((0, 2, 3)), (up * up).sum((0, 2, 3))])
    err = ((got - want).abs() / scale).max().item()
    assert err < 2e-5, err                                         # fp32 per-thread sums, fp64 across threads
