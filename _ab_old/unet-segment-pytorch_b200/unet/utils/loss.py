"""Segmentation losses on the fused statistics kernel — drop-in for ``unet.utils.loss``.

Same classes, constructor arguments, sub-module names and formulas as the reference
(reference: unet/utils/loss.py:18-271).  The per-pixel work (softmax, cross-entropy,
one-hot products and their sums) is one CUDA reduction pass (`ops.SegStats`, csrc/loss.cu)
producing four (N, C) tables; what remains of each loss is a handful of O(N*C) tensor
expressions written exactly as the reference combines them, differentiated by autograd,
whose gradient tables drive the single backward pass over the pixels.  No Python loop over
the batch, no boolean-mask indexing, no host synchronisation.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops


def _stats(predictions: torch.Tensor, targets: torch.Tensor):
    if predictions.dim() != 4:
        raise ValueError("predictions must be (N, C, H, W) logits")
    return ops.SegStats.apply(predictions, targets)


def _dice_from_stats(cnt, inter, psum, smooth, reduction, ignore_background):
    # loss.py:70-85
    dice = (2.0 * inter + smooth) / (psum + cnt + smooth)
    if ignore_background and dice.shape[1] > 1:
        dice = dice[:, 1:]
    if reduction == 'mean':
        return 1.0 - dice.mean()
    if reduction == 'sum':
        return (1.0 - dice).sum()
    return 1.0 - dice


def _balanced_ce_from_stats(cnt, ce, class_weight, smooth):
    # loss.py:134-148: class-1 pixels share `class_weight`, class-0 pixels share the rest
    n = cnt.shape[0]
    total = ce[:, 0] * ((1 - class_weight) / (cnt[:, 0] + smooth))
    if cnt.shape[1] > 1:
        total = total + ce[:, 1] * (class_weight / (cnt[:, 1] + smooth))
    return total.sum() / n


class DiceLoss(nn.Module):
    """Reference: unet/utils/loss.py:18-85."""

    def __init__(self, smooth: float = 1.0, reduction: str = 'mean', ignore_background: bool = True):
        super().__init__()
        self.smooth = smooth
        self.reduction = reduction
        self.ignore_background = ignore_background

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        cnt, _, inter, psum = _stats(predictions, targets)
        return _dice_from_stats(cnt, inter, psum, self.smooth, self.reduction, self.ignore_background)


class BalancedCELoss(nn.Module):
    """Reference: unet/utils/loss.py:88-150."""

    def __init__(self, class_weight: float = 0.5, smooth: float = 1e-6):
        super().__init__()
        self.class_weight = class_weight
        self.smooth = smooth

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        cnt, ce, _, _ = _stats(predictions, targets)
        return _balanced_ce_from_stats(cnt, ce, self.class_weight, self.smooth)


class DiceBCELoss(nn.Module):
    """Balanced CE + Dice from ONE statistics pass.  Reference: unet/utils/loss.py:153-191."""

    def __init__(self, ce_weight: float = 1.0, dice_weight: float = 1.0, class_weight: float = 0.5):
        super().__init__()
        self.ce_weight = ce_weight
        self.dice_weight = dice_weight
        self.balanced_ce = BalancedCELoss(class_weight=class_weight)
        self.dice_loss = DiceLoss(ignore_background=True)

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        cnt, ce, inter, psum = _stats(predictions, targets)
        bce = _balanced_ce_from_stats(cnt, ce, self.balanced_ce.class_weight, self.balanced_ce.smooth)
        d = self.dice_loss
        dice = _dice_from_stats(cnt, inter, psum, d.smooth, d.reduction, d.ignore_background)
        return self.ce_weight * bce + self.dice_weight * dice


class DeepSupervisionLoss(nn.Module):
    """Weighted sum of the base loss over [main, ds1, ds2, ds3].  Reference: loss.py:194-229."""

    def __init__(self, base_criterion: nn.Module, weights: list = None):
        super().__init__()
        self.base_criterion = base_criterion
        self.weights = weights or [1.0, 0.4, 0.2, 0.1]

    def forward(self, predictions, targets: torch.Tensor) -> torch.Tensor:
        if not isinstance(predictions, (list, tuple)):
            return self.base_criterion(predictions, targets)
        total = 0.0
        for pred, w in zip(predictions, self.weights):
            total = total + w * self.base_criterion(pred, targets)
        return total


def create_loss_function(loss_type: str = 'dice_bce', ce_weight: float = 1.0, dice_weight: float = 1.0,
                         class_weights: Optional[list] = None, balanced_class_weight: float = 0.5,
                         **kwargs) -> nn.Module:
    """Reference: unet/utils/loss.py:232-271."""
    kind = loss_type.lower()
    if kind == 'dice':
        return DiceLoss(ignore_background=True)
    if kind in ('ce', 'crossentropy'):
        weight = torch.tensor(class_weights, dtype=torch.float32) if class_weights is not None else None
        return nn.CrossEntropyLoss(weight=weight)
    if kind == 'balanced_ce':
        return BalancedCELoss(class_weight=balanced_class_weight)
    if kind == 'dice_bce':
        return DiceBCELoss(ce_weight=ce_weight, dice_weight=dice_weight, class_weight=balanced_class_weight)
    raise ValueError(f"Unknown loss type: {loss_type}")
