"""unet — B200-native drop-in for the reference's ``unet`` package (hot path only).

``unet.models.{UNet, AttentionUNet}``, the six layer blocks, ``unet.utils.loss`` and
``unet.utils.metrics`` keep the reference's API; the work underneath runs as
hand-written sm_100a kernels from ``libunetb200.so``.
"""
__version__ = "0.1.0"

from .fp32 import set_precision
from .models.layers import AttentionGate, AttentionUp, DoubleConv, Down, OutConv, Up
from .models.unet import AttentionUNet, UNet

__all__ = ["UNet", "AttentionUNet", "DoubleConv", "Down", "Up", "OutConv", "AttentionGate", "AttentionUp",
           "set_precision"]

# Everything else of the reference's package (unet.data.dataset, unet.utils.callbacks, unet.utils.plots, …)
# falls through to an attached reference checkout: see unet/overlay.py.
import sys as _sys

from . import overlay as _overlay

_overlay.install(_sys.modules[__name__])
__getattr__ = _overlay.package_getattr(__name__, ())
