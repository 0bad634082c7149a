"""GPU parity of the tcgen05 convolution kernels (through the C ABI) against a
plain fp32 torch convolution of the same bf16-rounded operands.

Tolerance: operands are identical bf16 values on both sides and accumulation is
fp32, so the only differences are summation order and the final bf16 rounding of
the output: |diff| <= 2^-7 * |ref| + small absolute slack.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(n, h, w, c, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)


def _ref_conv(xs, wt, taps):
    """xs: list of NHWC bf16 cpu tensors; wt: (Cout, Cin, k, k) bf16."""
    x = torch.cat([t.float() for t in xs], dim=3).permute(0, 3, 1, 2)
    y = F.conv2d(x, wt.float(), padding=1 if taps == 9 else 0)
    return y.permute(0, 2, 3, 1).contiguous()


def _pack(wt):
    cout, cin, k, _ = wt.shape
    return wt.permute(0, 2, 3, 1).reshape(cout, k * k, cin).contiguous()


def _close(got, ref, what):
    got = got.float().cpu()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad} / {ref.numel()} mismatches, max err {err.max().item():.4g}"


CASES = [
    # n, h, w, c0, c1, cout, taps
    (2, 16, 16, 64, 0, 64, 9),
    (1, 8, 8, 64, 0, 128, 9),
    (2, 16, 16, 128, 0, 64, 1),
    (1, 32, 32, 64, 64, 64, 9),     # virtual concat
    (1, 10, 12, 64, 0, 64, 9),      # odd sizes -> masked tiles
    (3, 4, 4, 128, 0, 512, 9),      # several images per tile, two N tiles
    (1, 8, 8, 32, 0, 32, 9),        # 64B swizzle path
    (1, 8, 8, 16, 16, 48, 9),       # 32B swizzle path, Cout not a multiple of 32
    (1, 16, 256, 64, 0, 64, 9),     # wide rows (BW = 128)
    (2, 32, 32, 512, 512, 512, 9),  # up1.0-like, long K loop
    # wide rows (W % 128 == 0): halo-resident kernel
    (2, 5, 128, 64, 64, 64, 9),     # two sources, H not a multiple of the row block
    (1, 7, 256, 128, 0, 128, 9),    # two channel chunks, N tile 128
    (1, 4, 128, 64, 0, 256, 9),     # two N tiles
    (3, 9, 128, 64, 0, 32, 9),
]


@pytest.mark.parametrize("n,h,w,c0,c1,cout,taps", CASES)
def test_conv_fwd(n, h, w, c0, c1, cout, taps):
    from unet import kernels as K

    k = 3 if taps == 9 else 1
    xs = [_mk(n, h, w, c0, 1)] + ([_mk(n, h, w, c1, 2)] if c1 else [])
    g = torch.Generator().manual_seed(3)
    wt = (torch.randn(cout, c0 + c1, k, k, generator=g) / (taps * (c0 + c1)) ** 0.5).to(torch.bfloat16)
    ref = _ref_conv(xs, wt, taps)
    dev = [t.cuda() for t in xs]
    out, st = K.conv_fwd(dev[0], _pack(wt).cuda(), taps, x1=dev[1] if c1 else None, stats=True)
    torch.cuda.synchronize()
    _close(out, ref, "conv_fwd")
    # statistics are of the stored (bf16-rounded) values
    o = out.double().cpu().reshape(-1, cout)
    s = st.sum(dim=0).cpu()
    assert torch.allclose(s[0], o.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[1], (o * o).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("w", [16, 128])
def test_conv_fwd_epilogue_options(w):
    from unet import kernels as K

    n, h, c, cout = 2, 16, 64, 128
    x = _mk(n, h, w, c, 5)
    g = torch.Generator().manual_seed(6)
    wt = (torch.randn(cout, c, 3, 3, generator=g) / 24).to(torch.bfloat16)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g)
    ref = torch.relu(_ref_conv([x], wt, 9) * scale + shift)
    out = K.conv_fwd(x.cuda(), _pack(wt).cuda(), 9, scale=scale.cuda(), shift=shift.cuda(), relu=True)
    _close(out, ref, "affine+relu")
    # accumulate + split output (the dgrad of a virtual concat)
    base0 = _mk(n, h, w, 64, 7)
    base1 = _mk(n, h, w, 64, 8)
    o0, o1 = base0.cuda().clone(), base1.cuda().clone()
    K.conv_fwd(x.cuda(), _pack(wt).cuda(), 9, out=o0, out1=o1, split=64, accumulate=True)
    r = _ref_conv([x], wt, 9)
    _close(o0, r[..., :64] + base0.float(), "split0+acc")
    _close(o1, r[..., 64:] + base1.float(), "split1+acc")


WG_CASES = [
    (2, 16, 16, 64, 0, 64, 9),
    (1, 8, 8, 64, 0, 128, 9),
    (2, 16, 16, 128, 0, 32, 1),
    (1, 32, 32, 64, 64, 64, 9),
    (1, 10, 12, 64, 0, 64, 9),
    (1, 8, 8, 32, 0, 32, 9),
    (1, 8, 8, 16, 16, 48, 9),
    (2, 32, 32, 256, 0, 512, 9),
    (4, 64, 64, 64, 0, 64, 9),
    # wide rows: halo-resident wgrad kernel
    (2, 5, 128, 64, 64, 64, 9),
    (1, 7, 256, 128, 0, 128, 9),    # two tap groups
    (3, 9, 128, 64, 0, 32, 9),
    (1, 4, 128, 64, 0, 96, 9),
    (2, 16, 512, 64, 0, 64, 9),
    (2, 6, 128, 128, 128, 128, 9),  # four channel chunks: two CTA pairs (cta_group::2 wgrad)
    (1, 12, 256, 192, 64, 64, 9),   # 64-byte-swizzled half of dy per CTA
    (2, 7, 128, 64, 0, 128, 9),     # single channel chunk: the CTA pair splits the taps
    (1, 5, 256, 64, 0, 64, 9),
    # deep layers: two-CTA per-tap kernel (Cout a multiple of 128)
    (1, 16, 16, 128, 0, 256, 9),    # 9 M tiles: odd, the last pair has a phantom tile
    (2, 8, 8, 256, 256, 128, 1),    # 1x1, two sources
    (4, 32, 32, 512, 0, 512, 9),    # down4-like: two N tiles
]


@pytest.mark.parametrize("n,h,w,c0,c1,cout,taps", WG_CASES)
def test_conv_wgrad(n, h, w, c0, c1, cout, taps):
    from unet import kernels as K

    k = 3 if taps == 9 else 1
    xs = [_mk(n, h, w, c0, 11)] + ([_mk(n, h, w, c1, 12)] if c1 else [])
    dy = _mk(n, h, w, cout, 13)
    x = torch.cat([t.float() for t in xs], dim=3).permute(0, 3, 1, 2).requires_grad_(False)
    wt = torch.zeros(cout, c0 + c1, k, k, requires_grad=True)
    y = F.conv2d(x, wt, padding=1 if taps == 9 else 0)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wt.grad  # (Cout, Cin, k, k)
    dev = [t.cuda() for t in xs]
    part = K.conv_wgrad(dev[0], dy.cuda(), taps, x1=dev[1] if c1 else None)
    got = part.sum(dim=0).cpu()  # (taps*Cin, Cout)
    got = got.reshape(k, k, c0 + c1, cout).permute(3, 2, 0, 1)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, f"wgrad max err {err} vs scale {scale}"


@pytest.mark.parametrize("cout,cin,taps,splits", [(64, 64, 9, 1), (64, 64, 9, 37), (32, 64, 1, 5), (48, 80, 9, 9),
                                                   (256, 512, 9, 2), (16, 16, 1, 12)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_wgrad_reduce(cout, cin, taps, splits, accumulate):
    """Split-K fold into the OIHW .grad: both the direct and the pre-summed (> 8 splits) path."""
    from unet import kernels as K

    k = 3 if taps == 9 else 1
    g = torch.Generator().manual_seed(5)
    part = torch.randn(splits, taps * cin, cout, generator=g)
    base = torch.randn(cout, cin, k, k, generator=g)
    ref = part.double().sum(0).reshape(k, k, cin, cout).permute(3, 2, 0, 1).float()
    if accumulate:
        ref = ref + base
    grad = base.clone().cuda()
    K.wgrad_reduce(part.cuda(), cout, cin, taps, grad, accumulate=accumulate)
    assert torch.allclose(grad.cpu(), ref, rtol=1e-5, atol=1e-5 * splits ** 0.5)


@pytest.mark.parametrize("cout,cin,k", [(64, 64, 3), (128, 64, 3), (32, 64, 1), (48, 80, 3), (512, 1024, 3), (16, 16, 1)])
def test_pack_conv_weight(cout, cin, k):
    """OIHW fp32 -> the two bf16 packs, bit-exact (round to nearest even), with and without a folded scale."""
    from unet import kernels as K

    g = torch.Generator().manual_seed(6)
    w = torch.randn(cout, cin, k, k, generator=g)
    sc = torch.rand(cout, generator=g) + 0.5
    for scale in (None, sc):
        ws = w if scale is None else w * scale.view(-1, 1, 1, 1)
        fwd, dg = K.pack_conv_weight(w.cuda(), True, True, out_scale=None if scale is None else scale.cuda())
        ref_f = ws.permute(0, 2, 3, 1).reshape(cout, k * k, cin).bfloat16()
        ref_d = ws.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, k * k, cout).bfloat16()
        assert torch.equal(fwd.cpu(), ref_f)
        assert torch.equal(dg.cpu(), ref_d)


def test_weight_packer_multi():
    """One launch packs every registered weight exactly like the per-tensor kernel."""
    from unet import kernels as K

    g = torch.Generator().manual_seed(8)
    ws = [torch.randn(s, generator=g).cuda() for s in [(64, 64, 3, 3), (128, 64, 3, 3), (32, 64, 1, 1),
                                                        (256, 512, 3, 3), (64, 1, 3, 3), (1, 32, 1, 1)]]
    packer = K.WeightPacker(ws)
    packer.run()
    assert packer.get(ws[4]) is None and packer.get(ws[5]) is None   # stem / psi: not tensor-core packs
    for w in ws[:4]:
        fwd, dg = packer.get(w)
        rf, rd = K.pack_conv_weight(w, True, True)
        assert torch.equal(fwd, rf) and torch.equal(dg, rd)
    assert packer.get(ws[0].clone()) is None


def test_wgrad_reduce_multi():
    """Several layers folded by one call == the per-layer fold (both split regimes, both tap counts)."""
    from unet import kernels as K

    g = torch.Generator().manual_seed(9)
    shapes = [(64, 64, 9, 37), (128, 64, 9, 4), (32, 64, 1, 12), (256, 512, 9, 2), (48, 80, 9, 9), (16, 16, 1, 1)]
    items, refs = [], []
    for cout, cin, taps, splits in shapes:
        k = 3 if taps == 9 else 1
        part = torch.randn(splits, taps * cin, cout, generator=g).cuda()
        base = torch.randn(cout, cin, k, k, generator=g).cuda()
        ref = K.wgrad_reduce(part.clone(), cout, cin, taps, base.clone(), accumulate=True)
        tgt = base.clone()
        items.append((part, cout, cin, taps, tgt))
        refs.append(ref)
    K.wgrad_reduce_multi(items, accumulate=True)
    for (_, _, _, _, tgt), ref in zip(items, refs):
        assert torch.equal(tgt, ref)
