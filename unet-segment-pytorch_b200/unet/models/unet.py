"""UNet / AttentionUNet on the fused B200 kernels — drop-in for ``unet.models.unet``.

Same constructor arguments, attribute and child names (hence ``state_dict`` layout),
``forward`` signature and return convention as the reference
(reference: unet/models/unet.py:16-217).  Input: any float ``(N, C, H, W)`` CUDA tensor;
output: fp32 ``(N, n_classes, H, W)`` logits (a list ``[main, ds1, ds2, ds3]`` in training
mode with deep supervision).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .layers import AttentionUp, DoubleConv, Down, OutConv, Up


class _UNetBase(nn.Module):
    def _encode(self, x):
        """Encoder: each stage's bandwidth pass also emits the 2x2 max-pooled tensor the next
        stage consumes, so MaxPool2d (layers.py:56) never runs as its own pass."""
        x1, p = self.inc._run(x, None, pool_out=True)
        x2, p = self.down1.forward_pooled(p, True)
        x3, p = self.down2.forward_pooled(p, True)
        x4, p = self.down3.forward_pooled(p, True)
        x5, _ = self.down4.forward_pooled(p, False)
        return x1, x2, x3, x4, x5

    def get_num_params(self, trainable_only: bool = True) -> int:
        params = self.parameters()
        return sum(p.numel() for p in params if p.requires_grad or not trainable_only)


class UNet(_UNetBase):
    """Reference: unet/models/unet.py:16-106."""

    def __init__(self, n_channels: int = 1, n_classes: int = 2, bilinear: bool = True,
                 base_features: int = 64):
        super().__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        bf = base_features
        factor = 2 if bilinear else 1
        self.inc = DoubleConv(n_channels, bf)
        self.down1 = Down(bf, bf * 2)
        self.down2 = Down(bf * 2, bf * 4)
        self.down3 = Down(bf * 4, bf * 8)
        self.down4 = Down(bf * 8, bf * 16 // factor)
        self.up1 = Up(bf * 16, bf * 8 // factor, bilinear)
        self.up2 = Up(bf * 8, bf * 4 // factor, bilinear)
        self.up3 = Up(bf * 4, bf * 2 // factor, bilinear)
        self.up4 = Up(bf * 2, bf, bilinear)
        self.outc = OutConv(bf, n_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x1, x2, x3, x4, x5 = self._encode(x)
        d = self.up1(x5, x4)
        d = self.up2(d, x3)
        d = self.up3(d, x2)
        d = self.up4(d, x1)
        return self.outc(d)


class AttentionUNet(_UNetBase):
    """Reference: unet/models/unet.py:109-217."""

    def __init__(self, n_channels: int = 1, n_classes: int = 2, bilinear: bool = True,
                 base_features: int = 64, deep_supervision: bool = False):
        super().__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        self.deep_supervision = deep_supervision
        bf = base_features
        factor = 2 if bilinear else 1
        self.inc = DoubleConv(n_channels, bf)
        self.down1 = Down(bf, bf * 2)
        self.down2 = Down(bf * 2, bf * 4)
        self.down3 = Down(bf * 4, bf * 8)
        self.down4 = Down(bf * 8, bf * 16 // factor)
        self.up1 = AttentionUp(bf * 16, bf * 8 // factor, bilinear)
        self.up2 = AttentionUp(bf * 8, bf * 4 // factor, bilinear)
        self.up3 = AttentionUp(bf * 4, bf * 2 // factor, bilinear)
        self.up4 = AttentionUp(bf * 2, bf, bilinear)
        self.outc = OutConv(bf, n_classes)
        if deep_supervision:
            # creation order ds_out3, ds_out2, ds_out1 as in the reference (unet.py:171-173)
            self.ds_out3 = OutConv(bf * 8 // factor, n_classes)
            self.ds_out2 = OutConv(bf * 4 // factor, n_classes)
            self.ds_out1 = OutConv(bf * 2 // factor, n_classes)

    def forward(self, x: torch.Tensor):
        size = x.shape[2:]
        x1, x2, x3, x4, x5 = self._encode(x)
        d4 = self.up1(x5, x4)
        d3 = self.up2(d4, x3)
        d2 = self.up3(d3, x2)
        d1 = self.up4(d2, x1)
        logits = self.outc(d1)
        if self.deep_supervision and self.training:
            # auxiliary heads (unet.py:204-209): 2-channel logits, resampled to the input size
            ds = [ops.resize_logits(head(d), size)
                  for head, d in ((self.ds_out1, d2), (self.ds_out2, d3), (self.ds_out3, d4))]
            return [logits] + ds
        return logits
