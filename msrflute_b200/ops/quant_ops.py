"""Segmented gradient quantization over a flat arena.

``quantize_segments_`` = the reference transform (``extensions/quantization/quant.py:9-50``) applied per
tensor segment.  CUDA path: ``csrc/misc_kernels.cu`` (segmented min/max kernel + fused encode kernel; the
|g| quantile of each segment comes from ``torch.quantile``).  ``pack_segments`` / ``unpack_add_`` are the
real wire format: ``bits``-wide level codes + a 1-bit/elem keep mask + (lo, hi) per segment.
"""
from typing import Sequence, Tuple

import torch

from . import _ext


def _seg_table(segments, device):
    if torch.is_tensor(segments):
        return segments.to(device=device, dtype=torch.int64).reshape(-1, 2)
    return torch.tensor(list(segments), dtype=torch.int64, device=device).reshape(-1, 2)


def segment_stats(flat: torch.Tensor, segments, quant_threshold: float, global_stats=False) -> torch.Tensor:
    """``[n_seg, 3]`` (lo, hi, thresh) per segment."""
    from ..extensions.quantization.quant import find_min_max_gradient
    segs = segments.tolist() if torch.is_tensor(segments) else list(segments)
    if global_stats:
        allv = torch.cat([flat[o:o + n] for o, n in segs])
        st = torch.stack(find_min_max_gradient(allv, quant_threshold))
        return st.unsqueeze(0).expand(len(segs), 3).contiguous()
    return torch.stack([torch.stack(find_min_max_gradient(flat[o:o + n], quant_threshold)) for o, n in segs])


def quantize_segments_(flat: torch.Tensor, segments, quant_bits: int, quant_threshold: float, global_stats=False,
                       return_stats: bool = False):
    """In-place simulated quantization.  CUDA (``csrc/misc_kernels.cu``): segmented min/max in one pass over the arena
    (ordered-int atomics), |g| quantiles per segment from ``torch.quantile`` (sort-based, exact like the reference),
    then ONE encode kernel over the whole arena driven by the segment table.  ``return_stats``: also return the
    ``[n_seg, 3]`` (lo, hi, thresh) table the levels were built from (what :func:`wire_encode` needs)."""
    if _ext.use_cuda_kernels(flat) and not global_stats and flat.is_contiguous() and flat.dtype == torch.float32:
        ext = _ext.load()
        tab = _seg_table(segments, flat.device).contiguous()
        stats = ext.seg_minmax(flat.view(-1), tab)
        segs = segments.tolist() if torch.is_tensor(segments) else list(segments)
        stats[:, 2] = torch.stack([_abs_quantile(flat[o:o + n], quant_threshold) for o, n in segs])
        ext.quantize_segments(flat.view(-1), tab, stats, int(quant_bits), False)
        _ext.count_launch(3)
        return (flat, stats) if return_stats else flat
    from ..extensions.quantization.quant import quantize_tensor_
    stats = segment_stats(flat, segments, quant_threshold, global_stats)
    segs = segments.tolist() if torch.is_tensor(segments) else list(segments)
    for i, (o, n) in enumerate(segs):
        quantize_tensor_(flat[o:o + n], quant_bits, quant_threshold, tuple(stats[i]))
    return (flat, stats) if return_stats else flat


# ------------------------------------------------------------------------------------------------ wire format
def wire_encode(values: torch.Tensor, sizes, lo_hi: torch.Tensor, quant_bits: int):
    """Pack an ALREADY quantised dense gradient (``values``: concatenation of tensors of ``sizes`` elements, each
    snapped to ``2**bits`` levels on its ``[lo, hi]`` or zeroed) into what travels between ranks:

        codes   uint8 level index per element (bits <= 8) or two uint8 byte planes (9..16 bits),
        bitmap  1 bit per element (1 = the element survived the magnitude threshold),
        table   float32 [n_seg, 2] = (lo, level width).

    ``4 N`` bytes become ``N + N/8`` (8-bit) or ``2N + N/8`` (9..16-bit).  Decoding reproduces ``lo + code * width``."""
    assert 1 <= quant_bits <= 16
    n_bins = 2 ** quant_bits
    dev = values.device
    sz = torch.as_tensor(list(sizes), dtype=torch.int64, device=dev)
    lo = lo_hi[:, 0].to(torch.float32)
    width = (lo_hi[:, 1].to(torch.float32) - lo) / (n_bins - 1)
    lo_e = torch.repeat_interleave(lo, sz)
    w_e = torch.repeat_interleave(torch.where(width > 0, width, torch.ones_like(width)), sz)
    v = values.reshape(-1).to(torch.float32)
    idx = torch.round((v - lo_e) / w_e).clamp_(0, n_bins - 1).to(torch.int32)
    if quant_bits <= 8:
        codes = idx.to(torch.uint8)
    else:                                        # two byte planes (every collective backend moves uint8)
        codes = torch.cat([(idx & 0xFF).to(torch.uint8), (idx >> 8).to(torch.uint8)])
    keep = v != 0
    pad = (-keep.numel()) % 8
    kb = torch.cat([keep, keep.new_zeros(pad)]).view(-1, 8).to(torch.uint8)
    bitmap = (kb * (2 ** torch.arange(8, device=dev, dtype=torch.uint8))).sum(dim=1).to(torch.uint8)
    return codes, bitmap, torch.stack([lo, width], dim=1).contiguous()


def wire_decode(codes: torch.Tensor, bitmap: torch.Tensor, table: torch.Tensor, sizes) -> torch.Tensor:
    """Inverse of :func:`wire_encode`: dense fp32 vector."""
    dev = codes.device
    sz = torch.as_tensor(list(sizes), dtype=torch.int64, device=dev)
    total = int(sz.sum())
    lo_e = torch.repeat_interleave(table[:, 0], sz)
    w_e = torch.repeat_interleave(table[:, 1], sz)
    bits = ((bitmap.unsqueeze(1) >> torch.arange(8, device=dev, dtype=torch.uint8)) & 1).bool().view(-1)[:total]
    if codes.numel() == 2 * total:
        idx = codes[:total].to(torch.float32) + 256.0 * codes[total:].to(torch.float32)
    else:
        idx = codes.to(torch.float32)
    vals = lo_e + idx * w_e
    return torch.where(bits, vals, torch.zeros_like(vals))


def _abs_quantile(x: torch.Tensor, q: float) -> torch.Tensor:
    a = x.reshape(-1).abs()
    if a.numel() > 2 ** 24:       # torch.quantile's input limit; kthvalue has none
        k = min(max(int(round(q * (a.numel() - 1))) + 1, 1), a.numel())
        return a.kthvalue(k).values
    return torch.quantile(a, q)


def pack_segments(flat: torch.Tensor, segments, quant_bits: int, quant_threshold: float):
    """Encode to (codes uint16/uint8, keep-mask uint8 bitmap, stats[n_seg,3]).  Lossless w.r.t.
    :func:`quantize_segments_` — ``unpack`` reproduces exactly the simulated-quantization values."""
    assert quant_bits <= 16
    stats = segment_stats(flat, segments, quant_threshold)
    segs = segments.tolist() if torch.is_tensor(segments) else list(segments)
    n_bins = 2 ** quant_bits
    code_dtype = torch.uint8 if quant_bits <= 8 else torch.int32
    codes = torch.zeros(flat.numel(), dtype=code_dtype, device=flat.device)
    keep = torch.zeros(flat.numel(), dtype=torch.bool, device=flat.device)
    for i, (o, n) in enumerate(segs):
        g = flat[o:o + n]
        lo, hi, th = stats[i]
        width = (hi - lo) / (n_bins - 1)
        safe = torch.where(width > 0, width, torch.ones_like(width))
        codes[o:o + n] = torch.ceil((g - lo) / safe - 0.5).clamp_(0, n_bins - 1).to(code_dtype)
        keep[o:o + n] = g.abs() > th
    pad = (-keep.numel()) % 8
    kb = torch.cat([keep, keep.new_zeros(pad)]).view(-1, 8).to(torch.uint8)
    bitmap = (kb * (2 ** torch.arange(8, device=flat.device, dtype=torch.uint8))).sum(dim=1).to(torch.uint8)
    return codes, bitmap, stats


def unpack_add_(out: torch.Tensor, codes, bitmap, stats, segments, quant_bits: int, alpha: float = 1.0):
    """``out += alpha · dequant(codes, bitmap)`` — the decode half, fused with the server-side accumulation."""
    segs = segments.tolist() if torch.is_tensor(segments) else list(segments)
    n_bins = 2 ** quant_bits
    bits = ((bitmap.unsqueeze(1) >> torch.arange(8, device=bitmap.device, dtype=torch.uint8)) & 1).bool().view(-1)
    keep = bits[:out.numel()]
    for i, (o, n) in enumerate(segs):
        lo, hi, _ = stats[i]
        width = (hi - lo) / (n_bins - 1)
        vals = lo + codes[o:o + n].to(out.dtype) * width
        out[o:o + n].add_(torch.where(keep[o:o + n], vals, torch.zeros_like(vals)), alpha=alpha)
    return out
