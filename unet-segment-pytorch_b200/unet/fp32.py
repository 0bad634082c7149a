"""fp32 / TF32 evaluation mode: ``unet.set_precision("tf32")``.

The default precision of the B200 path is bf16 operands with fp32 accumulation.  This module is
the higher-accuracy *inference* path of BASELINE configs[0] ("AttentionUNet fp32 forward") and of
north_star's "fp32/TF32 mode: logits within 1e-3 relative": activations stay fp32 (NHWC), the
convolutions run on the tensor cores as ``kind::tf32`` with operands rounded to TF32 where they are
produced, BatchNorm uses the running statistics (eval mode, folded into the conv epilogue) and the
stem, pooling, resampling, gate and output head are plain fp32 kernels (csrc/fp32_eval.cu).
Forward only — training runs in bf16.  Reference semantics: unet/models/layers.py:16-255.
"""
from __future__ import annotations

import torch

from . import _C
from ._C import ptr, stream
from . import kernels as K

F32 = torch.float32
PRECISION = "bf16"


def set_precision(mode: str) -> None:
    """"bf16" (default: training and inference) or "tf32" (fp32 activations, inference only)."""
    global PRECISION
    if mode not in ("bf16", "tf32", "fp32"):
        raise ValueError("precision must be 'bf16' or 'tf32'")
    PRECISION = "tf32" if mode in ("tf32", "fp32") else "bf16"


def active() -> bool:
    return PRECISION == "tf32"


def _check(module, *tensors):
    if module.training:
        raise NotImplementedError("TF32 mode is forward-only and runs BatchNorm with the running statistics: call "
                                  "model.eval(), or unet.set_precision('bf16') to train")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("unet-b200 modules run on CUDA tensors only (no CPU fallback)")


def nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW -> dense NHWC fp32 (no copy if already channels_last fp32)."""
    return x.float().permute(0, 2, 3, 1).contiguous()


def nchw(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 3, 1, 2)


def _bn_coeffs(bn):
    return K.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.eps))


def _pack(w):
    cout, cin, kh, kw = w.shape
    out = torch.empty((cout, kh * kw, cin), device=w.device, dtype=F32)
    _C.call("ub2_f32_pack_weight", ptr(w.contiguous()), ptr(out), cout, cin, kh * kw, stream())
    return out


def conv(x0, x1, weight, scale=None, shift=None, relu=False):
    """x0 / x1: NHWC fp32; returns NHWC fp32 (TF32 tensor-core convolution)."""
    n, h, w, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    cout, cin, kh, kw = weight.shape
    assert cin == c0 + c1
    out = torch.empty((n, h, w, cout), device=x0.device, dtype=F32)
    _C.call("ub2_conv_fwd_tf32", ptr(x0), c0, c0, ptr(x1), c1, c1, ptr(_pack(weight)), ptr(out), cout, n, h, w,
            cout, kh * kw, ptr(scale), ptr(shift), int(relu), stream())
    return out


def double_conv(module, x0, x1=None, pool_out=False):
    """DoubleConv (layers.py:16-41) on logical-NCHW inputs; returns (activation, pooled or None)."""
    _check(module, x0, x1)
    seq = module.double_conv
    cin = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
    s1, h1 = _bn_coeffs(seq[1])
    if cin % 16 != 0:
        if x1 is not None:
            raise RuntimeError("concatenated inputs need channel counts that are multiples of 16")
        x = x0.contiguous().float()
        n, _, h, w = x.shape
        cout = seq[0].weight.shape[0]
        a = torch.empty((n, h, w, cout), device=x.device, dtype=F32)
        _C.call("ub2_f32_conv_in", ptr(x), ptr(seq[0].weight.contiguous()), ptr(s1), ptr(h1), ptr(a), n, cin, h, w,
                cout, stream())
    else:
        a = conv(nhwc(x0), nhwc(x1) if x1 is not None else None, seq[0].weight, s1, h1, relu=True)
    s2, h2 = _bn_coeffs(seq[4])
    a = conv(a, None, seq[3].weight, s2, h2, relu=True)
    p = maxpool_nhwc(a) if pool_out else None
    return nchw(a), (nchw(p) if p is not None else None)


def maxpool_nhwc(a):
    n, h, w, c = a.shape
    out = torch.empty((n, h // 2, w // 2, c), device=a.device, dtype=F32)
    _C.call("ub2_f32_maxpool", ptr(a), ptr(out), n, h, w, c, stream())
    return out


def maxpool(x):
    return nchw(maxpool_nhwc(nhwc(x)))


def upsample(x, out_h, out_w):
    """nn.Upsample(2x, bilinear, align_corners=True) + F.pad to (out_h, out_w) (layers.py:78, :98-102)."""
    a = nhwc(x)
    n, h, w, c = a.shape
    out = torch.empty((n, out_h, out_w, c), device=a.device, dtype=F32)
    _C.call("ub2_f32_upsample", ptr(a), ptr(out), n, h, w, 2 * h, 2 * w, out_h, out_w, c, stream())
    return nchw(out)


def gate(module, g, x):
    """AttentionGate.forward (layers.py:171-192)."""
    _check(module, g, x)
    gn, xn = nhwc(g), nhwc(x)
    n, h, w, cx = xn.shape
    _, hin, win, _ = gn.shape
    q = conv(gn, None, module.W_g[0].weight)      # W_g at low resolution: commutes with the resampling
    xp = conv(xn, None, module.W_x[0].weight)
    sg, hg = _bn_coeffs(module.W_g[1])
    sx, hx = _bn_coeffs(module.W_x[1])
    sp, hp = _bn_coeffs(module.psi[1])
    ci = q.shape[3]
    out = torch.empty_like(xn)
    _C.call("ub2_f32_gate", ptr(q), ptr(xp), ptr(xn), ptr(sg), ptr(hg), ptr(sx), ptr(hx),
            ptr(module.psi[0].weight.reshape(-1).contiguous()), ptr(sp), ptr(hp), ptr(out), n, hin, win, h, w, ci,
            cx, stream())
    return nchw(out)


def outc(module, x):
    """OutConv (layers.py:109-123): fp32 NCHW logits."""
    _check(module, x)
    a = nhwc(x)
    n, h, w, c = a.shape
    wt = module.conv.weight
    k = wt.shape[0]
    logits = torch.empty((n, k, h, w), device=a.device, dtype=F32)
    _C.call("ub2_f32_outc", ptr(a), ptr(wt.reshape(k, -1).contiguous()), ptr(module.conv.bias), ptr(logits), n, h, w,
            c, k, stream())
    return logits
