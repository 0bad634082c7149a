"""U-Net model components (same exports as the reference's ``unet.models``)."""
from .layers import AttentionGate, AttentionUp, DoubleConv, Down, OutConv, Up
from .unet import AttentionUNet, UNet

__all__ = ["DoubleConv", "Down", "Up", "OutConv", "AttentionGate", "AttentionUp", "UNet", "AttentionUNet"]

import sys as _sys

from .. import overlay as _overlay

_overlay.install(_sys.modules[__name__])
