"""Normalisation layers fused with their epilogues.

``group_norm``: statistics over (channels-in-group × spatial) per sample; affine either per channel
(``torch.nn.GroupNorm``) or **per group** (the reference's ``GroupNorm2d``,
``experiments/cv_resnet_fedcifar100/group_normalization.py:59-84``); optional residual add and ReLU in the same
pass.  ``layer_norm``: last-dim LayerNorm (+ optional residual).  CUDA: ``csrc/norm_kernels.cu`` — one CTA per
(sample, group) row, Welford in registers + shuffle reduction, the backward recomputes x̂ from the saved
(mean, rstd) and emits dx plus per-row dγ/dβ partials that a tiny second kernel folds.
"""
import torch
import torch.nn.functional as F

from . import _ext


def _group_norm_ref(x, num_groups, weight, bias, eps, residual, relu, per_group_affine):
    N, C = x.shape[0], x.shape[1]
    xf = x.float()
    xg = xf.reshape(N, num_groups, -1)
    mean = xg.mean(dim=2, keepdim=True)
    var = xg.var(dim=2, unbiased=False, keepdim=True)
    y = (xg - mean) * torch.rsqrt(var + eps)
    if weight is not None:
        if per_group_affine:
            y = y * weight.float().view(1, num_groups, 1) + bias.float().view(1, num_groups, 1)
            y = y.reshape(x.shape)
        else:
            y = y.reshape(x.shape)
            shape = (1, C) + (1,) * (x.dim() - 2)
            y = y * weight.float().view(shape) + bias.float().view(shape)
    else:
        y = y.reshape(x.shape)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = F.relu(y)
    return y.to(x.dtype)


class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, num_groups, eps, relu, per_group_affine):
        ext = _ext.load()
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        y, mean, rstd = ext.group_norm_fwd(x, weight, bias, res, int(num_groups), float(eps), bool(relu),
                                           bool(per_group_affine))
        _ext.count_launch(1)
        ctx.save_for_backward(x, weight, mean, rstd, y if relu else None)
        ctx.cfg = (int(num_groups), bool(relu), bool(per_group_affine), residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd, y = ctx.saved_tensors
        G, relu, pga, has_res = ctx.cfg
        ext = _ext.load()
        dx, dw, db, dres = ext.group_norm_bwd(dy.contiguous(), x, weight, mean, rstd, y, G, relu, pga, has_res)
        _ext.count_launch(2)
        return dx, dw, db, (dres if has_res else None), None, None, None, None


def group_norm(x, num_groups, weight=None, bias=None, eps=1e-5, residual=None, relu=False, per_group_affine=False):
    if x.is_cuda and weight is not None and _ext.load() is not None and x.dtype in (torch.float32, torch.bfloat16) \
            and hasattr(_ext.load(), "group_norm_fwd"):
        return _GroupNormFn.apply(x, weight, bias, residual, num_groups, eps, relu, per_group_affine)
    return _group_norm_ref(x, num_groups, weight, bias, eps, residual, relu, per_group_affine)


def layer_norm(x, weight, bias, eps=1e-5, residual=None):
    """LayerNorm over the last dim (+ residual added *before* normalisation, the transformer pattern)."""
    if residual is not None:
        x = x + residual
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)
