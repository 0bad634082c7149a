"""Segmentation metrics on a device-side confusion matrix — drop-in for ``unet.utils.metrics``.

Reference: unet/utils/metrics.py:16-227.  ``update`` no longer moves tensors to the host or
loops over pixels: one kernel (csrc/metrics.cu) accumulates an int64 (C+1)x(C+1) histogram on
the GPU; the host sees it only when ``confusion_matrix`` / ``compute`` is read.  Counts are
integers, so results are bit-identical to the reference on identical predictions.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .. import kernels as K


def _accumulate(predictions, targets, num_classes, cm, ignore_index=None, threshold=None):
    if not (predictions.is_cuda and targets.is_cuda):
        raise RuntimeError("unet-b200 metrics run on CUDA tensors only (no CPU fallback)")
    if predictions.dim() == 4:
        predictions = predictions.detach().contiguous().float()
    else:
        predictions = predictions.detach().long()
    return K.confusion(predictions, targets.detach().long(), num_classes, cm, ignore_index=ignore_index,
                       threshold=threshold)


class SegmentationMetrics:
    """Reference: unet/utils/metrics.py:16-157 (same public attributes and methods)."""

    def __init__(self, num_classes: int = 2, class_names: Optional[List[str]] = None,
                 ignore_index: Optional[int] = None):
        self.num_classes = num_classes
        self.class_names = class_names or [f'class_{i}' for i in range(num_classes)]
        self.ignore_index = ignore_index
        self._dev_cm = None                      # (C+1, C+1) int64 on the GPU, created lazily
        self._host_cm = np.zeros((num_classes, num_classes), dtype=np.int64)

    # the reference exposes a plain numpy attribute (metrics.py:47); keep it readable/writable
    @property
    def confusion_matrix(self) -> np.ndarray:
        if self._dev_cm is not None:
            c = self.num_classes
            self._host_cm = self._host_cm + self._dev_cm[:c, :c].cpu().numpy()
            self._dev_cm.zero_()
        return self._host_cm

    @confusion_matrix.setter
    def confusion_matrix(self, value) -> None:
        self._host_cm = np.asarray(value, dtype=np.int64)
        if self._dev_cm is not None:
            self._dev_cm.zero_()

    def reset(self) -> None:
        self._host_cm = np.zeros((self.num_classes, self.num_classes), dtype=np.int64)
        if self._dev_cm is not None:
            self._dev_cm.zero_()

    def update(self, predictions: torch.Tensor, targets: torch.Tensor, threshold: Optional[float] = None) -> None:
        """predictions: (N,C,H,W) logits or (N,H,W) class indices; targets (N,H,W).
        ``threshold`` (extension): for 2-class logits use softmax[:,1] > threshold
        (scripts/predict.py:155-159) instead of argmax."""
        if self._dev_cm is None or self._dev_cm.device != predictions.device:
            _ = self.confusion_matrix  # flush counts held on another device
            self._dev_cm = torch.zeros((self.num_classes + 1, self.num_classes + 1), dtype=torch.int64,
                                       device=predictions.device)
        _accumulate(predictions, targets, self.num_classes, self._dev_cm, self.ignore_index, threshold)

    def compute(self) -> Dict[str, float]:
        cm = self.confusion_matrix
        total = cm.sum()
        if total == 0:
            return self._empty_results()
        class_iou, class_dice = {}, {}
        for i, name in enumerate(self.class_names[:self.num_classes]):
            tp = cm[i, i]
            fp = cm[:, i].sum() - tp
            fn = cm[i, :].sum() - tp
            class_iou[name] = tp / (tp + fp + fn) if (tp + fp + fn) > 0 else 0.0
            class_dice[name] = 2 * tp / (2 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else 0.0
        ious = [v for v in class_iou.values() if v > 0]
        dices = [v for v in class_dice.values() if v > 0]
        return {
            'pixel_accuracy': float(np.diag(cm).sum() / total),
            'mean_iou': float(np.mean(ious)) if ious else 0.0,
            'mean_dice': float(np.mean(dices)) if dices else 0.0,
            'class_iou': class_iou,
            'class_dice': class_dice,
        }

    def _empty_results(self) -> Dict[str, float]:
        return {
            'pixel_accuracy': 0.0, 'mean_iou': 0.0, 'mean_dice': 0.0,
            'class_iou': {name: 0.0 for name in self.class_names},
            'class_dice': {name: 0.0 for name in self.class_names},
        }

    def get_confusion_matrix(self) -> np.ndarray:
        return self.confusion_matrix.copy()


def _full_histogram(predictions, targets, num_classes):
    cm = torch.zeros((num_classes + 1, num_classes + 1), dtype=torch.int64, device=predictions.device)
    return _accumulate(predictions, targets, num_classes, cm).float()


def compute_iou(predictions: torch.Tensor, targets: torch.Tensor, num_classes: int = 2,
                smooth: float = 1e-6) -> torch.Tensor:
    """Per-class IoU; reference: unet/utils/metrics.py:160-192."""
    cm = _full_histogram(predictions, targets, num_classes)
    inter = cm.diagonal()[:num_classes]
    union = cm.sum(0)[:num_classes] + cm.sum(1)[:num_classes] - inter
    return (inter + smooth) / (union + smooth)


def compute_dice(predictions: torch.Tensor, targets: torch.Tensor, num_classes: int = 2,
                 smooth: float = 1e-6) -> torch.Tensor:
    """Per-class Dice; reference: unet/utils/metrics.py:195-227."""
    cm = _full_histogram(predictions, targets, num_classes)
    inter = cm.diagonal()[:num_classes]
    return (2.0 * inter + smooth) / (cm.sum(0)[:num_classes] + cm.sum(1)[:num_classes] + smooth)
