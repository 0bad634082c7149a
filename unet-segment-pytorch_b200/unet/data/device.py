"""uint8 slices -> normalised fp32 images + int64 targets on the GPU, double buffered.

Reference behaviour restated here (and nowhere on the host):
  * ``LungTumorDataset.__getitem__`` (unet/data/dataset.py:146-171): image = L-mode PNG / 255,
    mask = (label PNG > 127) as int64;
  * ``apply_basic_transforms`` (unet/data/augmentations.py:117-170), the transform the reference
    uses when albumentations is absent: a second uint8 round trip of the image (:148; the identity on
    all 256 grey levels in float32), random
    horizontal flip of image and mask for training (:160-162), ``(image - mean) / std`` (:165),
    image ``(1,H,W)`` float, mask ``(H,W)`` long;
  * ``preprocess_image`` (scripts/predict.py:100-136): the same without the flip.
Resizing is the host's business (the slices the converter writes are already model-sized, and PIL's
``resize`` to the same size is a copy).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .. import kernels as K


def prepare_batch(images: torch.Tensor, labels: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None,
                  mean: float = 0.5, std: float = 0.5):
    """``images`` uint8 ``(N,H,W)`` or ``(N,1,H,W)`` and ``labels`` uint8 ``(N,H,W)`` on a CUDA device ->
    ``(x fp32 (N,1,H,W), targets int64 (N,H,W) or None)``.  ``flags`` uint8 ``(N)``: bit 0 flips an
    image and its label horizontally, bit 1 vertically.  Bit-exact with ``apply_basic_transforms`` and
    with ``predict.py``'s ``preprocess_image`` (the former's second uint8 round trip is the identity)."""
    if not images.is_cuda:
        raise RuntimeError("unet-b200 ops need CUDA tensors: there is no CPU fallback")
    if images.dim() == 4:
        if images.shape[1] != 1:
            raise ValueError("prepare_batch handles single-channel slices (n_channels=1)")
        images = images[:, 0]
    return K.prepare_batch(images.contiguous(), None if labels is None else labels.contiguous(), flags, mean, std)


def draw_flags(n: int, generator: torch.Generator, hflip_prob: float, vflip_prob: float) -> torch.Tensor:
    """Per-image flip bits (host side, uint8): bit 0 horizontal with probability ``hflip_prob``
    (the reference flips when ``np.random.rand() > 0.5``, augmentations.py:160), bit 1 vertical."""
    u = torch.rand((2, n), generator=generator)
    return ((u[0] < hflip_prob).to(torch.uint8) | ((u[1] < vflip_prob).to(torch.uint8) << 1)).contiguous()


class _Slot:
    def __init__(self):
        self.host = {}      # name -> pinned staging tensor
        self.dev = {}       # name -> device staging tensor
        self.out = None     # (x, targets)
        self.ready = None   # recorded on the copy stream when `out` is complete


class DeviceBatchPipeline:
    """Iterate a host loader and hand out device batches one step ahead of the consumer.

    ``loader`` yields ``(images, labels)`` (or ``(images, labels, flags)``) per batch, either
      * uint8 slices ``(N,H,W)`` / ``(N,1,H,W)`` and uint8 labels ``(N,H,W)`` — expanded on the GPU
        by ``ub2_prepare_batch`` (flip bits drawn here when ``hflip_prob`` / ``vflip_prob`` > 0), or
      * what the reference's own ``DataLoader`` yields, fp32 ``(N,1,H,W)`` and int64 ``(N,H,W)`` —
        passed through (asynchronous copy only).
    Batches are staged in pinned memory and copied on a private stream while the previous step runs;
    ``__next__`` makes the caller's current stream wait for the batch and returns device tensors that
    stay valid until ``depth`` more batches have been requested.
    """

    def __init__(self, loader: Iterable, device, mean: float = 0.5, std: float = 0.5,
                 hflip_prob: float = 0.0, vflip_prob: float = 0.0, seed: int = 0, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceBatchPipeline stages batches for a CUDA device (no CPU fallback)")
        if depth < 2:
            raise ValueError("depth must be at least 2 (one batch in use, one in flight)")
        self.loader = loader
        self.mean, self.std = float(mean), float(std)
        self.hflip_prob, self.vflip_prob = float(hflip_prob), float(vflip_prob)
        self._gen = torch.Generator().manual_seed(seed)
        self._slots = [_Slot() for _ in range(depth)]
        self._stream = None
        self._it = None
        self._pending = None
        self._next_slot = 0

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        self._it = iter(self.loader)
        self._pending = self._issue()
        return self

    def __next__(self):
        slot = self._pending
        if slot is None:
            raise StopIteration
        self._pending = self._issue()     # the batch after this one goes in flight first
        current = torch.cuda.current_stream(self.device)
        current.wait_event(slot.ready)
        for t in slot.out:
            if t is not None:
                t.record_stream(current)   # allocated under the copy stream, read on this one
        return slot.out

    # ------------------------------------------------------------------ internals
    def _stage(self, slot: _Slot, name: str, t: torch.Tensor) -> torch.Tensor:
        """host tensor -> pinned staging -> device staging (asynchronous, on the copy stream)."""
        t = t.contiguous()
        h = slot.host.get(name)
        if h is None or h.shape != t.shape or h.dtype != t.dtype:
            h = slot.host[name] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            slot.dev[name] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        h.copy_(t)
        d = slot.dev[name]
        d.copy_(h, non_blocking=True)
        return d

    def _issue(self) -> Optional[_Slot]:
        try:
            batch = next(self._it)
        except StopIteration:
            return None
        images, labels = batch[0], batch[1]
        flags = batch[2] if len(batch) > 2 else None
        slot = self._slots[self._next_slot]
        self._next_slot = (self._next_slot + 1) % len(self._slots)
        if slot.ready is not None:
            slot.ready.synchronize()      # its previous copies have left the pinned buffers
        else:
            slot.ready = torch.cuda.Event()
        # whatever the consumer queued on its stream so far may still read this slot's tensors
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            if images.dtype == torch.uint8:
                if images.dim() == 4:
                    images = images[:, 0]
                n = images.shape[0]
                if flags is None and (self.hflip_prob > 0.0 or self.vflip_prob > 0.0):
                    flags = draw_flags(n, self._gen, self.hflip_prob, self.vflip_prob)
                d_img = self._stage(slot, "images", images)
                d_lab = self._stage(slot, "labels", labels.to(torch.uint8)) if labels is not None else None
                d_flg = self._stage(slot, "flags", flags.to(torch.uint8)) if flags is not None else None
                x = slot.dev.get("x")
                if x is None or x.shape != (n, 1) + tuple(images.shape[1:]):
                    x = slot.dev["x"] = torch.empty((n, 1) + tuple(images.shape[1:]), dtype=torch.float32,
                                                    device=self.device)
                    slot.dev["t"] = torch.empty(tuple(images.shape), dtype=torch.int64, device=self.device)
                t = slot.dev["t"] if d_lab is not None else None
                K.prepare_batch(d_img, d_lab, d_flg, self.mean, self.std, x=x, targets=t)
                slot.out = (x, t)
            else:
                slot.out = (self._stage(slot, "x", images), self._stage(slot, "t", labels) if labels is not None else None)
            slot.ready.record(self._stream)
        return slot
