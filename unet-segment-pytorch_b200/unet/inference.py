"""Inference engine: the eval-mode forward of a (Attention)UNet with the weight packs built once
and the whole forward replayed as a CUDA graph.

The reference predicts one slice at a time (scripts/predict.py:138-240: ``model(x)`` under
``torch.no_grad()``, softmax, threshold) and validates batch by batch (scripts/train.py:164-197).
At batch 1-4 a 512^2 forward is ~130 kernel launches and the host needs ~3 ms to issue what the GPU
executes in well under a millisecond; replaying a captured graph removes that, and packing the
bf16 weights once (they do not change at inference time) removes 25 launches per forward.
"""
from __future__ import annotations

import torch

from . import kernels as K
from . import ops
from .kernels import WeightPacker


class InferenceEngine:
    """``engine = InferenceEngine(model); logits = engine(x)`` — same result as ``model.eval()(x)``.

    The returned tensor is a static buffer that the next call with the same input shape overwrites
    (clone it to keep it).  Call ``refresh()`` after changing the model's parameters or buffers
    (e.g. ``load_state_dict``)."""

    def __init__(self, model: torch.nn.Module, cuda_graph: bool = True, warmup: int = 2):
        self.model = model.eval()
        self.cuda_graph = cuda_graph
        self.warmup = warmup
        self._graphs = {}
        self._eager_runs = {}
        self._packer = None
        self.refresh()

    def refresh(self) -> None:
        weights = [m.weight for m in self.model.modules() if isinstance(m, torch.nn.Conv2d) and m.weight.is_cuda]
        self._packer = WeightPacker(weights) if weights else None
        if self._packer is not None:
            self._packer.run()
        self._graphs.clear()
        self._eager_runs.clear()

    def _forward(self, x):
        prev, ops.PACKS = ops.PACKS, self._packer
        try:
            return self.model(x)
        finally:
            ops.PACKS = prev

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.model.training:
            raise RuntimeError("InferenceEngine runs the eval-mode forward: do not switch the model to train()")
        dev = next(self.model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("unet-b200 modules run on CUDA tensors only (no CPU fallback)")
        if not self.cuda_graph:
            return self._forward(x.to(dev, non_blocking=True))
        key = (tuple(x.shape), x.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            done = self._eager_runs.get(key, 0)
            if done < self.warmup:   # lazy initialisation / allocator warm-up on a side stream
                self._eager_runs[key] = done + 1
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    out = self._forward(x.to(dev, non_blocking=True))
                torch.cuda.current_stream(dev).wait_stream(side)
                return out
            gx = torch.empty(x.shape, dtype=x.dtype, device=dev)
            gx.copy_(x, non_blocking=True)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gout = self._forward(gx)
            entry = self._graphs[key] = (graph, gx, gout)
        graph, gx, gout = entry
        gx.copy_(x, non_blocking=True)
        graph.replay()
        return gout

    @torch.no_grad()
    def predict(self, x: torch.Tensor, threshold: float = 0.5, mean: float = 0.5, std: float = 0.5):
        """``predict_single`` (scripts/predict.py:204-240) for a batch, without leaving the GPU.

        ``x``: normalised fp32 ``(N,1,H,W)`` (``preprocess_image``'s tensor) or raw uint8 slices
        ``(N,H,W)`` / ``(N,1,H,W)``, which are normalised on the device as ``preprocess_image``
        (predict.py:126-127) does.  Returns ``(mask, tumor_ratio)``: ``mask`` uint8 ``(N,H,W)`` with
        255 where ``softmax(logits)[1] > threshold`` (``postprocess_mask``, predict.py:155-159) and
        ``tumor_ratio`` fp64 ``(N)`` = set pixels / pixels (predict.py:238).  Both are device tensors;
        nothing synchronises."""
        dev = next(self.model.parameters()).device
        x = x.to(dev, non_blocking=True)
        if x.dtype == torch.uint8:
            if x.dim() == 4:
                x = x[:, 0]
            x, _ = K.prepare_batch(x.contiguous(), None, None, mean, std)
        logits = self(x)
        if isinstance(logits, (list, tuple)):
            logits = logits[0]
        mask, positives = K.predict_mask(logits.contiguous(), threshold)
        return mask, positives.to(torch.float64) / float(mask.shape[1] * mask.shape[2])
