"""Shared model plumbing: classifier wrapper implementing the ``BaseModel`` contract once.

The reference re-implements ``loss``/``inference`` in every task file, each doing its own ``.to(device)`` and an
``.item()`` per batch (e.g. ``experiments/cv_lr_mnist/model.py:24-36``).  ``ClassifierModel`` does one forward
per call, keeps accuracy on the device (``loss_and_metrics`` serves evaluation with a single forward — SURVEY K23)
and runs the network under bf16 autocast on GPUs when ``compute_dtype: bf16`` (fp32 master weights in the arena).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..core.model import BaseModel


class ClassifierModel(BaseModel):
    """``net(x) -> logits``; CE loss; accuracy metric."""

    ignore_index = -100
    class_dim = 1            # logits layout (N, C, ...) like F.cross_entropy expects

    def __init__(self, model_config=None):
        super().__init__()
        cfg = model_config if model_config is not None else {}
        self.compute_dtype = str(cfg.get("compute_dtype", "fp32")) if hasattr(cfg, "get") else "fp32"

    # -- helpers -------------------------------------------------------------
    def _dev(self):
        return next(self.parameters()).device

    def _xy(self, batch):
        dev = self._dev()
        x, y = batch["x"], batch["y"]
        if x.device != dev:
            x = x.to(dev, non_blocking=True)
        if y.device != dev:
            y = y.to(dev, non_blocking=True)
        return x, y.long()

    def forward(self, x):
        if self.compute_dtype == "bf16" and x.is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.net(x).float()
        return self.net(x)

    def _accuracy(self, logits, y):
        pred = torch.argmax(logits, dim=self.class_dim)
        if self.ignore_index >= 0:
            mask = y != self.ignore_index
            return ((pred == y) & mask).sum().float() / mask.sum().clamp(min=1)
        return (pred == y).float().mean()

    # -- BaseModel contract -----------------------------------------------------
    def loss(self, input):
        x, y = self._xy(input)
        return F.cross_entropy(self.forward(x), y, ignore_index=self.ignore_index)

    def inference(self, input):
        x, y = self._xy(input)
        logits = self.forward(x)
        return {"output": logits, "acc": self._accuracy(logits, y).item(), "batch_size": x.shape[0]}

    def loss_and_metrics(self, input):
        x, y = self._xy(input)
        logits = self.forward(x)
        loss = F.cross_entropy(logits, y, ignore_index=self.ignore_index)
        return loss, {"output": None, "acc": self._accuracy(logits, y), "batch_size": x.shape[0]}
