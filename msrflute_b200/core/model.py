"""Abstract model interface every task plug-in implements (ref. ``core/model.py:7-51``).

``loss(batch) -> scalar tensor`` and ``inference(batch) -> dict`` with at least
``output``, ``acc``, ``batch_size``; extra metrics are either plain numbers
(treated as higher-is-better) or ``{'value':…, 'higher_is_better':…}``.

B200 additions (all optional, used by the device-resident engine):

* ``loss_and_metrics(batch)`` — ONE forward producing loss and metrics on the
  device (the reference runs two forwards per eval batch).
* ``device_batch(batch)``    — move/cast a collated batch onto this rank's GPU.
"""
from abc import ABC, abstractmethod

import torch


class BaseModel(ABC, torch.nn.Module):
    @abstractmethod
    def __init__(self, **kwargs):
        super().__init__()

    @abstractmethod
    def loss(self, input):
        """Forward pass + loss; returns a scalar tensor."""

    @abstractmethod
    def inference(self, input):
        """Forward pass + metrics; returns a dict (see module docstring)."""

    def set_eval(self):
        self.eval()

    def set_train(self):
        self.train()

    # ---- optional fast paths -------------------------------------------
    def device_batch(self, batch):
        dev = next(self.parameters()).device
        if isinstance(batch, dict):
            return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
        return batch

    def loss_and_metrics(self, batch):
        """Default: fall back to the two-call protocol."""
        loss = self.loss(batch)
        return loss, self.inference(batch)
