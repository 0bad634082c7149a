"""Device side of the input pipeline.

The reference's ``unet.data`` (``LungTumorDataset``, the albumentations pipelines) decodes PNG slices
on the host and is used unchanged; what moves to the GPU is the part of
``LungTumorDataset.__getitem__`` that runs per pixel on every slice — scale, normalise, binarise the
label, flip — and the host->device hand-over, so that 2 bytes per pixel cross PCIe instead of 12.
"""
from .device import DeviceBatchPipeline, prepare_batch

__all__ = ["DeviceBatchPipeline", "prepare_batch"]
