"""U-Net model components (same exports as the reference's ``unet.models``)."""
from .layers import AttentionGate, AttentionUp, DoubleConv, Down, OutConv, Up
from .unet import AttentionUNet, UNet

__all__ = ["DoubleConv", "Down", "Up", "OutConv", "AttentionGate", "AttentionUp", "UNet", "AttentionUNet"]
