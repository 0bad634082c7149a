"""Typed Python wrappers over the C ABI (one function per entry point family).

All tensors are CUDA tensors; activations are NHWC bf16 ``(N, H, W, C)`` and
must be contiguous in their last dimension.  These wrappers only allocate
outputs / workspaces with torch and pass raw pointers through ctypes.
"""
from __future__ import annotations

import torch

from . import _C
from ._C import byref, c_double, c_float, c_int, c_longlong, ptr, stream

BF16 = torch.bfloat16


def num_sms() -> int:
    return int(_C.lib().ub2_num_sms())


def _nhwc(t: torch.Tensor):
    assert t.dim() == 4 and t.dtype == BF16 and t.stride(3) == 1, "expected NHWC bf16"
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "expected dense NHWC rows"
    return n, h, w, c, ld


# --------------------------------------------------------------------------- convolution
def conv_fwd(x0, wgt, taps, x1=None, out=None, out1=None, split=0, scale=None, shift=None,
             relu=False, accumulate=False, stats=False, bn_override=0, grid_override=0):
    """Implicit-GEMM conv (3x3 pad 1 or 1x1). ``wgt``: (Cout, taps, C0+C1) bf16.

    Returns ``out`` or ``(out, stats_partials)`` where stats_partials is a
    (rows, 2, Cout) float64 tensor of per-CTA sums / sums of squares.
    """
    n, h, w, c0, ld0_in = _nhwc(x0)
    c1, ld1_in = 0, 0
    if x1 is not None:
        n1, h1, w1, c1, ld1_in = _nhwc(x1)
        assert (n1, h1, w1) == (n, h, w)
    cout = wgt.shape[0]
    assert wgt.dtype == BF16 and wgt.is_contiguous() and wgt.numel() == cout * taps * (c0 + c1)
    if out is None:
        out = torch.empty((n, h, w, cout if out1 is None else split), device=x0.device, dtype=BF16)
    _, _, _, _, ldo0 = _nhwc(out)
    ldo1 = 0
    if out1 is not None:
        _, _, _, _, ldo1 = _nhwc(out1)
    st = None
    rows = 0
    if stats:
        rows = num_sms()
        st = torch.empty((rows, 2, cout), device=x0.device, dtype=torch.float64)
    used = c_int(0)
    _C.call("ub2_conv_fwd", ptr(x0), ld0_in, c0, ptr(x1), ld1_in, c1, ptr(wgt), ptr(out), ldo0,
            ptr(out1), ldo1, split, n, h, w, cout, taps, ptr(scale), ptr(shift), int(relu),
            int(accumulate), ptr(st), rows, byref(used), bn_override, grid_override, stream())
    if stats:
        return out, st[: used.value]
    return out


def conv_wgrad(x0, dy, taps, x1=None, splits_override=0):
    """Split-K partial weight gradients: (splits, taps*(C0+C1), Cout) fp32."""
    n, h, w, c0, ld0_in = _nhwc(x0)
    c1, ld1_in = 0, 0
    if x1 is not None:
        _, _, _, c1, ld1_in = _nhwc(x1)
    n2, h2, w2, cout, ld_dy = _nhwc(dy)
    assert (n2, h2, w2) == (n, h, w)
    mtot = taps * (c0 + c1)
    max_splits = max(1, min(2 * num_sms(), (64 << 20) // (mtot * cout * 4)))
    partial = torch.empty((max_splits, mtot, cout), device=x0.device, dtype=torch.float32)
    used = c_int(0)
    _C.call("ub2_conv_wgrad", ptr(x0), ld0_in, c0, ptr(x1), ld1_in, c1, ptr(dy), ld_dy,
            ptr(partial), max_splits, byref(used), n, h, w, cout, taps, splits_override, stream())
    return partial[: used.value]
