"""Normalisation layers fused with their epilogues.

``group_norm``: statistics over (channels-in-group × spatial) per sample; affine either per channel
(``torch.nn.GroupNorm``) or **per group** (the reference's ``GroupNorm2d``,
``experiments/cv_resnet_fedcifar100/group_normalization.py:59-84``); optional residual add and ReLU in the same
pass.  CUDA: ``csrc/norm_kernels.cu`` — one warp per (sample, group) row, shuffle reductions, the backward
recomputes x̂ from the saved (mean, rstd) and emits dx, d_residual and per-row dγ/dβ partials.

Both autograd Functions carry a ``vmap`` rule so the device engine can run S simulated clients through ONE kernel
(``torch.func.vmap`` over the client dimension): the S per-client affine sets are passed as ``[S, A]`` and the kernel
picks the set from the row index — no per-client launches.
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


def _front(t, dim, S):
    """Move a (possibly unbatched) vmapped operand to a leading batch dim of size S."""
    if t is None:
        return None
    if dim is None:
        return t.unsqueeze(0).expand((S,) + tuple(t.shape))
    return t.movedim(dim, 0)


class _GNBwd(torch.autograd.Function):
    """dx, dγ, dβ, d_residual.  Not differentiable again (no double backward)."""

    @staticmethod
    def forward(dy, x, weight, mean, rstd, y, num_groups, relu, pga, has_res, sets):
        ext = _ext.load()
        dx, dw, db, dres = ext.group_norm_bwd(dy.contiguous(), x, weight, mean, rstd, y, num_groups, relu, pga,
                                              has_res, sets)
        _ext.count_launch(1)
        return dx, dw, db, (dres if has_res else dx.new_zeros(()))

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):
        raise RuntimeError("group_norm: double backward is not implemented")

    @staticmethod
    def vmap(info, in_dims, dy, x, weight, mean, rstd, y, num_groups, relu, pga, has_res, sets):
        S = info.batch_size
        dyb, xb = _front(dy, in_dims[0], S), _front(x, in_dims[1], S)
        wb = _front(weight, in_dims[2], S)
        mb, rb = _front(mean, in_dims[3], S), _front(rstd, in_dims[4], S)
        yb = _front(y, in_dims[5], S) if y is not None else None
        n = xb.shape[1]
        flat = lambda t: t.reshape((S * n,) + tuple(t.shape[2:])).contiguous()
        dx, dw, db, dres = _GNBwd.apply(flat(dyb), flat(xb), wb.reshape(S * sets, -1).contiguous(),
                                        mb.reshape(-1).contiguous(), rb.reshape(-1).contiguous(),
                                        flat(yb) if yb is not None else None, num_groups, relu, pga, has_res, S * sets)
        dx = dx.view((S, n) + tuple(dx.shape[1:]))
        dw, db = dw.view((S,) + tuple(weight.shape[-1:]) if sets == 1 else (S, sets, -1)), \
            db.view((S,) + tuple(weight.shape[-1:]) if sets == 1 else (S, sets, -1))
        if has_res:
            dres = dres.view((S, n) + tuple(dres.shape[1:]))
            return (dx, dw, db, dres), (0, 0, 0, 0)
        return (dx, dw, db, dres), (0, 0, 0, None)


class _GNFwd(torch.autograd.Function):
    @staticmethod
    def forward(x, weight, bias, residual, num_groups, eps, relu, pga, sets):
        ext = _ext.load()
        y, mean, rstd = ext.group_norm_fwd(x.contiguous(), weight, bias,
                                           residual.contiguous() if residual is not None else None, num_groups, eps,
                                           relu, pga, sets)
        _ext.count_launch(1)
        return y, mean, rstd

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, weight, bias, residual, num_groups, eps, relu, pga, sets = inputs
        y, mean, rstd = output
        ctx.mark_non_differentiable(mean, rstd)
        ctx.save_for_backward(x, weight, mean, rstd, y if relu else None)
        ctx.cfg = (int(num_groups), bool(relu), bool(pga), residual is not None, int(sets))

    @staticmethod
    def backward(ctx, dy, _dmean, _drstd):
        x, weight, mean, rstd, y = ctx.saved_tensors
        G, relu, pga, has_res, sets = ctx.cfg
        dx, dw, db, dres = _GNBwd.apply(dy, x, weight, mean, rstd, y, G, relu, pga, has_res, sets)
        return dx, dw, db, (dres if has_res else None), None, None, None, None, None

    @staticmethod
    def vmap(info, in_dims, x, weight, bias, residual, num_groups, eps, relu, pga, sets):
        S = info.batch_size
        xb = _front(x, in_dims[0], S)
        wb, bb = _front(weight, in_dims[1], S), _front(bias, in_dims[2], S)
        rb = _front(residual, in_dims[3], S) if residual is not None else None
        n = xb.shape[1]
        flat = lambda t: t.reshape((S * n,) + tuple(t.shape[2:])).contiguous()
        y, mean, rstd = _GNFwd.apply(flat(xb), wb.reshape(S * sets, -1).contiguous(), bb.reshape(S * sets, -1).contiguous(),
                                     flat(rb) if rb is not None else None, num_groups, eps, relu, pga, S * sets)
        return (y.view((S, n) + tuple(y.shape[1:])), mean.view(S, -1), rstd.view(S, -1)), (0, 0, 0)


#: tests set this to compare against the plain PyTorch formulation on a GPU
FORCE_REFERENCE = False


def _cuda_gn_ok(x, weight):
    if FORCE_REFERENCE:
        return False
    if not x.is_cuda or weight is None or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    ext = _ext.load()
    return ext is not None and hasattr(ext, "group_norm_fwd")


def group_norm(x, num_groups, weight=None, bias=None, eps=1e-5, residual=None, relu=False, per_group_affine=False):
    if _cuda_gn_ok(x, weight):
        if residual is not None and residual.dtype != x.dtype:
            residual = residual.to(x.dtype)
        return _GNFwd.apply(x, weight, bias, residual, int(num_groups), float(eps), bool(relu), bool(per_group_affine), 1)[0]
    return _group_norm_ref(x, num_groups, weight, bias, eps, residual, relu, per_group_affine)


def layer_norm(x, weight, bias, eps=1e-5, residual=None):
    """LayerNorm over the last dim (+ residual added *before* normalisation, the transformer pattern)."""
    if residual is not None:
        x = x + residual
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


# ------------------------------------------------------------------------------------------------ LayerNorm
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ext = _ext.load()
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        w = weight.float().contiguous() if weight is not None else None
        b = bias.float().contiguous() if bias is not None else None
        y, mean, rstd = ext.layer_norm_fwd(x2, w, b, float(eps))
        _ext.count_launch(1)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.in_shape, ctx.affine = shp, weight is not None
        ctx.w_dtype = weight.dtype if weight is not None else None
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        ext = _ext.load()
        dx, dw, db = ext.layer_norm_bwd(dy.reshape(x2.shape).to(x2.dtype).contiguous(), x2, w, mean, rstd)
        _ext.count_launch(1)
        if not ctx.affine:
            return dx.view(ctx.in_shape), None, None, None
        return dx.view(ctx.in_shape), dw.to(ctx.w_dtype), db.to(ctx.w_dtype), None


def layer_norm(x, weight=None, bias=None, eps: float = 1e-5):
    """LayerNorm over the last dimension.  CUDA (fp32 / bf16): ``csrc/layernorm_kernels.cu`` — one warp per row, the
    backward produces dx, dγ and dβ in one kernel.  Elsewhere: ``F.layer_norm``."""
    if _ext.use_cuda_kernels(x) and x.dtype in (torch.float32, torch.bfloat16) and x.shape[-1] <= 6144 \
            and (weight is None) == (bias is None):
        return _LayerNormFn.apply(x, weight, bias, eps)
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


class FusedLayerNorm(torch.nn.LayerNorm):
    """Drop-in ``nn.LayerNorm`` (1-D normalized_shape) running :func:`layer_norm`."""

    def forward(self, x):
        if len(self.normalized_shape) != 1:
            return super().forward(x)
        return layer_norm(x, self.weight, self.bias, self.eps)


def swap_layer_norm_modules(module: torch.nn.Module) -> int:
    """Replace every 1-D ``nn.LayerNorm`` by a :class:`FusedLayerNorm` sharing its parameters; returns the count."""
    n = 0
    for name, child in list(module.named_children()):
        if type(child) is torch.nn.LayerNorm and len(child.normalized_shape) == 1:
            new = FusedLayerNorm(child.normalized_shape, eps=child.eps, elementwise_affine=child.elementwise_affine,
                                 device=child.weight.device if child.weight is not None else None)
            new.weight, new.bias = child.weight, child.bias
            setattr(module, name, new)
            n += 1
        else:
            n += swap_layer_norm_modules(child)
    return n
