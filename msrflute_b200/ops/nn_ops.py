"""Layer ops backed by ``csrc/nn_kernels.cu`` (SURVEY K3 BatchNorm, K6 embedding, K8 dropout); every op falls back to
the equivalent PyTorch expression on CPU — that expression is also the tests' oracle."""
import itertools

import torch
import torch.nn.functional as F

from . import _ext


def _cuda_ok(*ts):
    if not all(t is None or t.is_cuda for t in ts):
        return False
    ext = _ext.load()
    return ext is not None and hasattr(ext, "embedding_fwd")


# ------------------------------------------------------------------------------------------------ embedding (K6)
class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, weight, padding_idx):
        out = _ext.load().embedding_fwd(idx, weight)
        _ext.count_launch(1)
        ctx.save_for_backward(idx)
        ctx.V, ctx.padding_idx = weight.shape[0], -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dw = _ext.load().embedding_bwd(idx, dy.contiguous().float(), ctx.V, ctx.padding_idx)
        _ext.count_launch(1)
        return None, dw, None


def embedding(idx, weight, padding_idx=None):
    """Row gather; the backward is a scatter-add of gradient rows into a fresh ``[V, D]`` buffer (padding row skipped)."""
    if _cuda_ok(idx, weight) and weight.dtype == torch.float32 and weight.is_contiguous() and idx.dtype == torch.int64:
        return _EmbeddingFn.apply(idx, weight, padding_idx)
    return F.embedding(idx, weight, padding_idx=padding_idx)


class Embedding(torch.nn.Embedding):
    """``nn.Embedding`` whose CUDA forward / backward are this repo's kernels (same parameters, same state dict)."""

    def forward(self, idx):
        if self.max_norm is None and not self.sparse and not self.scale_grad_by_freq:
            return embedding(idx, self.weight, self.padding_idx)
        return super().forward(idx)


# ------------------------------------------------------------------------------------------------ dropout (K8)
_seed_counter = itertools.count(1)


class _DropoutMask(torch.autograd.Function):
    """``y = x * keep(seed) / (1 - p)`` — linear and self-adjoint, so the backward is the same op on ``dy`` with the same
    seed.  Written in the functorch-compatible style (``setup_context`` + ``vmap``) so the device engine's
    ``vmap(grad_and_value(...))`` wave step can run models that contain it."""

    @staticmethod
    def forward(x, p, seed):
        y = _ext.load().dropout_apply(x.contiguous(), float(p), seed, False)
        _ext.count_launch(1)
        return y.view_as(x)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.p = float(inputs[1])
        ctx.save_for_backward(inputs[2])

    @staticmethod
    def backward(ctx, dy):
        (seed,) = ctx.saved_tensors
        return _DropoutMask.apply(dy, ctx.p, seed), None, None

    @staticmethod
    def vmap(info, in_dims, x, p, seed):
        # elementwise over the flattened tensor: the batched tensor is just a bigger tensor (every element has its own
        # Philox counter); forward and backward both see the batch dimension in front, so the masks line up
        xd = in_dims[0]
        if xd is None:
            return _DropoutMask.apply(x, p, seed), None
        return _DropoutMask.apply(x.movedim(xd, 0).contiguous(), p, seed), 0


class Dropout(torch.nn.Module):
    """Philox dropout with the mask recomputed in the backward (no mask tensor).  The seed lives in a device counter
    that is advanced by a device op, so a CUDA graph that captured this layer draws a new mask on every replay."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = float(p)
        self._seed = None

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        if not (_cuda_ok(x) and x.dtype == torch.float32):
            return F.dropout(x, self.p, True)
        if self._seed is None or self._seed.device != x.device:
            base = (torch.initial_seed() * 0x9E3779B1 + next(_seed_counter) * 0x85EBCA77) & 0x3FFFFFFFFFFFFFFF
            self._seed = torch.tensor([base], dtype=torch.int64, device=x.device)
        used = self._seed.clone()                # the value THIS call uses; the counter moves on below
        y = _DropoutMask.apply(x, self.p, used)
        self._seed.add_(0x632BE5AB)               # device-side increment (captured by CUDA graphs)
        return y

    def extra_repr(self):
        return "p={}".format(self.p)


# ------------------------------------------------------------------------------------------------ BatchNorm2d (K3)
class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, run_mean, run_var, momentum, eps, relu):
        y, stats = _ext.load().batch_norm_fwd(x, gamma, beta, residual, run_mean, run_var, float(momentum), float(eps), bool(relu))
        _ext.count_launch(1)
        ctx.save_for_backward(x, y, gamma, stats)
        ctx.relu, ctx.has_res = bool(relu), residual is not None
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, y, gamma, stats = ctx.saved_tensors
        outs = _ext.load().batch_norm_bwd(dy.contiguous(), x, y, gamma, stats, ctx.relu, ctx.has_res)
        _ext.count_launch(1)
        dx, dgamma, dbeta = outs[0], outs[1], outs[2]
        dres = outs[3] if ctx.has_res else None
        return dx, (dgamma if gamma is not None else None), dbeta, dres, None, None, None, None, None


def batch_norm_train(x, gamma, beta, run_mean=None, run_var=None, momentum=0.1, eps=1e-5, residual=None, relu=False):
    """Training-mode BatchNorm2d (+ residual) (+ ReLU) on NCHW fp32; updates the running statistics in place."""
    if _cuda_ok(x, gamma, beta, residual) and x.dtype == torch.float32 and x.dim() == 4:
        return _BatchNormFn.apply(x, gamma, beta, residual, run_mean, run_var, momentum, eps, relu)[0]
    y = F.batch_norm(x, run_mean, run_var, gamma, beta, True, momentum, eps)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y
