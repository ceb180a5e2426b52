"""Autograd wrappers around the slot-batched kernels (``csrc/conv_kernels.cu``, ``csrc/norm_kernels.cu``).

"Slot-batched" = one launch processes the same layer of S simulated clients whose parameters are rows of the
``[S, P]`` arenas.  Parameters are addressed as (arena, offset) — they never appear as autograd leaves — and their
gradients are *accumulated by the backward kernels directly into the gradient arena*, which is exactly where the
fused clip/statistics/SGD kernel reads them.  Only activations flow through autograd; a dummy leaf keeps the graph
alive for layers whose input does not require grad (the stem convolution).
"""
import torch

import os

from . import _ext

_CONV_IMPLS = {"auto": 0, "fma": 1, "tcgen05": 2}
_impl_applied = False


def set_conv_impl(name: str) -> None:
    """Select the slot-batched convolution kernels: ``tcgen05`` (kind::tf32 implicit GEMM, tensor cores), ``fma``
    (exact fp32 CUDA-core tiles) or ``auto`` (tcgen05 whenever a slot's GEMM has >= 48 rows).  ``FLUTE_CONV_IMPL``
    sets the initial choice."""
    global _impl_applied
    _ext.load().slot_conv_set_impl(_CONV_IMPLS[name])
    _impl_applied = True


def _ensure_impl():
    if not _impl_applied:
        set_conv_impl(os.environ.get("FLUTE_CONV_IMPL", "auto"))


def live_taps(Hi, Wi, KH, KW, stride, pad):
    """Filter taps (kh, kw) that touch real input for at least one output position — same rule and order as
    ``set_taps`` in ``csrc/conv_kernels.cu``.  The others only ever multiply zero padding."""
    Ho, Wo = (Hi + 2 * pad - KH) // stride + 1, (Wi + 2 * pad - KW) // stride + 1
    rows = [kh for kh in range(KH) if any(0 <= oh * stride - pad + kh < Hi for oh in range(Ho))]
    cols = [kw for kw in range(KW) if any(0 <= ow * stride - pad + kw < Wi for ow in range(Wo))]
    return [(kh, kw) for kh in rows for kw in cols]


class SlotConv2d(torch.autograd.Function):
    """x [S, B, Cin, H, W] (fp32) * per-slot weight [Cout, Cin, KH, KW] at W[s, w_off:] → [S, B, Cout, Ho, Wo].
    ``compact``: the slot arena stores only the live taps of this filter, ``[Cout, Cin, len(live_taps)]``."""

    @staticmethod
    def forward(ctx, x, dummy, W, G, w_off, Cout, KH, KW, stride, pad, compact=False):
        ext = _ext.load()
        _ensure_impl()
        x = x.contiguous()
        y = ext.slot_conv_fprop(x, W, w_off, Cout, KH, KW, stride, pad, bool(compact))
        _ext.count_launch(1)
        ctx.save_for_backward(x)
        ctx.W, ctx.G, ctx.cfg = W, G, (w_off, KH, KW, stride, pad, bool(compact))
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w_off, KH, KW, stride, pad, compact = ctx.cfg
        ext = _ext.load()
        dy = dy.contiguous()
        ext.slot_conv_wgrad(x, dy, ctx.G, w_off, KH, KW, stride, pad, compact)    # accumulates into the grad arena
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ext.slot_conv_dgrad(dy, ctx.W, w_off, x.shape[2], x.shape[3], x.shape[4], KH, KW, stride, pad, compact)
            _ext.count_launch(1)
        _ext.count_launch(1)
        return dx, None, None, None, None, None, None, None, None, None, None


class SlotGroupNorm(torch.autograd.Function):
    """GroupNorm (+residual, +ReLU) on [S*B, C, H, W] with per-slot affine read from / grads written to the arenas."""

    @staticmethod
    def forward(ctx, x, residual, dummy, W, G, w_off, b_off, groups, eps, relu, per_group, sets):
        ext = _ext.load()
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, mean, rstd = ext.group_norm_fwd_arena(x, W, w_off, b_off, res, groups, eps, relu, per_group, sets)
        _ext.count_launch(1)
        ctx.save_for_backward(x, mean, rstd, y if relu else None)
        ctx.W, ctx.G = W, G
        ctx.cfg = (w_off, b_off, groups, relu, per_group, sets, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, y = ctx.saved_tensors
        w_off, b_off, groups, relu, per_group, sets, has_res = ctx.cfg
        dx, dres = _ext.load().group_norm_bwd_arena(dy.contiguous(), x, ctx.W, w_off, mean, rstd, y, groups, relu,
                                                    per_group, has_res, sets, ctx.G, w_off, b_off)
        _ext.count_launch(1)
        return dx, (dres if has_res else None), None, None, None, None, None, None, None, None, None, None


class SlotLinear(torch.autograd.Function):
    """y[s] = x[s] · W[s]ᵀ + b[s] with W[s] = arena[s, w_off:].view(N, K); batched GEMMs, grads added into the arena."""

    @staticmethod
    def forward(ctx, x, dummy, W, G, w_off, b_off, N, K):
        S = x.shape[0]
        w = W[:, w_off:w_off + N * K].view(S, N, K)
        b = W[:, b_off:b_off + N].view(S, 1, N)
        ctx.save_for_backward(x)
        ctx.W, ctx.G, ctx.cfg = W, G, (w_off, b_off, N, K)
        return torch.baddbmm(b, x, w.transpose(1, 2))

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w_off, b_off, N, K = ctx.cfg
        S = x.shape[0]
        w = ctx.W[:, w_off:w_off + N * K].view(S, N, K)
        ctx.G[:, w_off:w_off + N * K].view(S, N, K).add_(torch.bmm(dy.transpose(1, 2), x))
        ctx.G[:, b_off:b_off + N].add_(dy.sum(dim=1))
        dx = torch.bmm(dy, w) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, None, None, None
