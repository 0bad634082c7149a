"""Device side of the input pipeline.

The reference's ``unet.data`` (``LungTumorDataset``, the albumentations pipelines) decodes PNG slices
on the host and is used unchanged; what moves to the GPU is the part of
``LungTumorDataset.__getitem__`` that runs per pixel on every slice — scale, normalise, binarise the
label, flip — and the host->device hand-over, so that 2 bytes per pixel cross PCIe instead of 12.
"""
from .device import DeviceBatchPipeline, prepare_batch

__all__ = ["DeviceBatchPipeline", "prepare_batch"]

# LungTumorDataset, get_train_transforms, get_val_transforms (the reference's unet/data/__init__.py
# exports) and the submodules dataset / augmentations come from an attached reference checkout
import sys as _sys

from .. import overlay as _overlay

_overlay.install(_sys.modules[__name__])
__getattr__ = _overlay.package_getattr(__name__, ("dataset", "augmentations"))
