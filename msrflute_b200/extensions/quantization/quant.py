"""Gradient quantization + sparsification (ref. ``extensions/quantization/quant.py``).

Reference semantics (``quant.py:9-100``), per tensor (or globally with ``global_stats``):
``lo, hi = min(g), max(g)``; ``thresh = quantile(|g|, quant_threshold)``; every value is snapped to the
nearest of ``2**quant_bits`` equally spaced levels in ``[lo, hi]`` (``bucketize(g − w/2, linspace)``); then
values with ``|g| <= thresh`` are zeroed.  The reference keeps the result in fp32 (simulated quantization).

Here the same transform is closed-form (``level = clamp(ceil((g−lo)/w − 1/2), 0, L−1)``; no ``linspace`` /
``bucketize`` tables) and runs over the flat gradient arena with a per-tensor segment table; on CUDA it is
two hand-written kernels (``csrc/misc_kernels.cu``: segmented min/max, then one encode kernel for the whole
arena).  Between ranks a quantised payload travels PACKED: ``quant_model`` leaves the per-tensor level table on the
gradient arena, ``make_payload`` attaches it and ``core.federated._send_gradients`` ships ``quant_bits``-wide level codes +
a 1-bit keep mask + (lo, width) per tensor (``ops.quant_ops.wire_encode`` / ``wire_decode``) instead of fp32 values.
"""
import logging
from typing import Optional, Tuple

import torch

from ...utils import print_rank


def find_min_max_gradient(gradient: torch.Tensor, quant_threshold: Optional[float] = None) -> Tuple:
    """(min, max, |g|-quantile threshold) of a tensor — device scalars, no host sync."""
    flat = gradient.reshape(-1)
    lo, hi = torch.aminmax(flat)
    a = flat.abs()
    if a.numel() > 2 ** 24:       # torch.quantile's input limit; kthvalue has none
        k = min(max(int(round(quant_threshold * (a.numel() - 1))) + 1, 1), a.numel())
        thresh = a.kthvalue(k).values
    else:
        thresh = torch.quantile(a, quant_threshold)
    return lo, hi, thresh


def quant_bins(gradients: torch.Tensor, n_bins: int, min_grad, max_grad) -> torch.Tensor:
    """Snap to ``n_bins`` uniform levels on [min_grad, max_grad] (ties resolve like ``bucketize`` − ½ bin)."""
    lo = torch.as_tensor(min_grad, dtype=gradients.dtype, device=gradients.device)
    hi = torch.as_tensor(max_grad, dtype=gradients.dtype, device=gradients.device)
    width = (hi - lo) / (n_bins - 1)
    safe = torch.where(width > 0, width, torch.ones_like(width))
    idx = torch.ceil((gradients - lo) / safe - 0.5).clamp_(0, n_bins - 1)
    return torch.where(width > 0, lo + idx * width, lo.expand_as(gradients))


def quantize_tensor_(g: torch.Tensor, quant_bits: int, quant_threshold: float, stats=None):
    lo, hi, thresh = stats if stats is not None else find_min_max_gradient(g, quant_threshold)
    binned = quant_bins(g, 2 ** quant_bits, lo, hi)
    g.copy_(torch.where(g.abs() > thresh, binned, torch.zeros_like(g)))
    return g


def quant_flat_(flat: torch.Tensor, segments, quant_bits: int, quant_threshold: float, global_stats=False,
                return_stats: bool = False):
    """Quantize a flat gradient buffer in place, tensor by tensor (``segments`` = [(offset, size), …])."""
    from ...ops import quant_ops
    return quant_ops.quantize_segments_(flat, segments, quant_bits, quant_threshold, global_stats, return_stats=return_stats)


def quant_model(model: torch.nn.Module, quant_bits: int = 8, quant_threshold: Optional[float] = None,
                global_stats: bool = False):
    """Quantize ``p.grad`` of every parameter in place; no-op when ``quant_threshold`` is None."""
    if quant_threshold is None:
        return
    print_rank("Performing Gradient Quantization with Prob. Threshold: {}".format(quant_threshold),
               loglevel=logging.DEBUG)
    from ...core.strategies.utils import grad_arena
    ga = grad_arena(model)
    if ga is not None:
        _, stats = quant_flat_(ga.flat, list(zip(ga.layout.offsets, ga.layout.sizes)), quant_bits, quant_threshold,
                               global_stats, return_stats=True)
        # level table of this payload: lets the transport ship level codes + a keep bitmap instead of fp32 values
        # (ops.quant_ops.wire_encode); consumed (and cleared) by strategies.fedavg.make_payload
        ga.quant_wire = {"bits": int(quant_bits), "lo_hi": stats[:, :2].detach().clone()} if quant_bits <= 16 else None
        return
    params = [p for p in model.parameters() if p.grad is not None]
    stats = None
    if global_stats:
        stats = find_min_max_gradient(torch.cat([p.grad.reshape(-1) for p in params]), quant_threshold)
    for p in params:
        quantize_tensor_(p.grad.data, quant_bits, quant_threshold, stats)
