"""Helpers shared by the aggregation strategies (ref. ``core/strategies/utils.py``)."""
import logging

import numpy as np
import torch

from ...parallel.arena import module_arena
from ...utils import print_rank, to_device


def filter_weight(weight):
    """NaN/inf → 0, cap at 100 (ref. ``strategies/utils.py:11-19``)."""
    if np.isnan(weight) or not np.isfinite(weight):
        return 0.0
    return 100 if weight > 100 else weight


def filter_weight_tensor(w: torch.Tensor) -> torch.Tensor:
    """Device version of :func:`filter_weight` for the sync-free path."""
    w = torch.where(torch.isfinite(w), w, torch.zeros_like(w))
    return w.clamp(max=100.0)


def grad_arena(model):
    """The flat gradient buffer of an arena-backed model, or None."""
    ar = module_arena(model)
    return None if ar is None or ar[1] is None else ar[1]


def payload_flat(payload, like: torch.Tensor = None):
    """Return the payload's gradients as one flat tensor if it carries one."""
    flat = payload.get("flat", None)
    if flat is not None and like is not None and flat.numel() == like.numel():
        return flat
    return None


def aggregate_gradients_inplace(model, gradients, flat=None):
    """``model.grad += gradients`` — one flat add when both sides are arena-backed, else per tensor
    (ref. ``strategies/utils.py:21-33``)."""
    ga = grad_arena(model)
    if ga is not None and flat is not None and flat.numel() == ga.flat.numel():
        ga.flat.add_(flat.to(ga.flat.device, non_blocking=True))
        return
    for p, g in zip(model.parameters(), gradients):
        g = g.to(p.device)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.add_(g)


def scale_gradients(model, scale):
    ga = grad_arena(model)
    if ga is not None:
        ga.flat.mul_(scale)
    else:
        for p in model.parameters():
            if p.grad is not None:
                p.grad.mul_(scale)
