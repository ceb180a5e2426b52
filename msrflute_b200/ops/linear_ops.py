"""Linear layers on the tcgen05 tensor cores.

``linear(x, weight, bias)`` = ``x @ weightᵀ + bias`` with bf16 operands / fp32 accumulation through the hand-written
TMA + tcgen05 GEMM (``csrc/gemm_tcgen05.cu``).  All three GEMMs of a training step run on it:

    forward   Y[M,N]  = X[M,K]  · W[N,K]ᵀ                    (both operands already K-major)
    dgrad     dX[M,K] = dY[M,N] · W[N,K]                     (W is reduction-major here: MN-major B operand)
    wgrad     dW[N,K] = dY[M,N]ᵀ · X[M,K]                    (both reduction-major: MN-major A and B operands)

The backward GEMMs read dY, X and W exactly as they sit in memory through MN-major UMMA descriptors
(``gemm_bf16_mn``) — no ``.t().contiguous()`` copies (three HBM round trips per layer per step in round 1).

Bias add (and an optional ReLU) are fused in the GEMM epilogue.  ``TCLinear`` is a drop-in ``nn.Linear`` replacement
and ``swap_linear_modules`` retrofits an existing model (e.g. a HuggingFace encoder).  Off-GPU, or for shapes TMA
cannot describe (K not a multiple of 8), it falls back to ``F.linear``.
"""
import torch
import torch.nn.functional as F

from . import _ext


def _gemm(a, b, bias=None, relu=False, out_fp32=False):
    out = _ext.load().gemm_bf16_tn(a, b, bias, relu, out_fp32)
    _ext.count_launch(1)
    return out


def _gemm_mn(a, b, a_mn, b_mn, out_fp32=False):
    out = _ext.load().gemm_bf16_mn(a, b, bool(a_mn), bool(b_mn), bool(out_fp32))
    _ext.count_launch(1)
    return out


def _mn_available():
    import os
    ext = _ext.load()
    return ext is not None and hasattr(ext, "gemm_bf16_mn") and os.environ.get("FLUTE_GEMM_MN", "1") == "1"


def tc_available(x, weight):
    if not (x.is_cuda and weight.is_cuda) or weight.shape[1] % 8 != 0:
        return False
    ext = _ext.load()
    return ext is not None and hasattr(ext, "gemm_bf16_tn")


class _TCLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
        w16 = weight.to(torch.bfloat16).contiguous()
        y = _gemm(x2, w16, bias, relu, out_fp32=False)
        ctx.save_for_backward(x2, w16, y if relu else None)
        ctx.relu, ctx.has_bias, ctx.in_shape, ctx.w_dtype = relu, bias is not None, shp, weight.dtype
        return y.view(shp[:-1] + (weight.shape[0],)).to(x.dtype if x.dtype != torch.float32 else torch.bfloat16)

    @staticmethod
    def backward(ctx, dy):
        x2, w16, y = ctx.saved_tensors
        N = w16.shape[0]
        dy2 = dy.reshape(-1, N).to(torch.bfloat16)
        if ctx.relu:
            dy2 = dy2 * (y > 0)
        dy2 = dy2.contiguous()
        dx = dw = db = None
        M, K = x2.shape
        mn = _mn_available() and N % 8 == 0 and K % 8 == 0
        if ctx.needs_input_grad[0]:
            if mn and K > 64:
                # dX = dY · W: W [N, K] is stored reduction-major for this product -> MN-major B operand, read in place
                dx = _gemm_mn(dy2, w16, False, True, out_fp32=False).view(ctx.in_shape)
            elif N % 8 == 0:
                dx = _gemm(dy2, w16.t().contiguous(), None, False, out_fp32=False).view(ctx.in_shape)
            else:
                dx = (dy2 @ w16).view(ctx.in_shape)
        if ctx.needs_input_grad[1]:
            if mn and K > 64:
                # dW = dYᵀ · X: dY [M, N] and X [M, K] are both stored reduction-major -> MN-major A and B, no copies
                dw = _gemm_mn(dy2, x2, True, True, out_fp32=True).to(ctx.w_dtype)
            elif M % 8 == 0:
                dw = _gemm(dy2.t().contiguous(), x2.t().contiguous(), None, False, out_fp32=True).to(ctx.w_dtype)
            else:
                dw = (dy2.t().float() @ x2.float()).to(ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.float().sum(dim=0)
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False):
    if tc_available(x, weight) and x.dtype in (torch.float32, torch.bfloat16):
        return _TCLinearFn.apply(x, weight, bias, relu)
    y = F.linear(x, weight, bias)
    return F.relu(y) if relu else y


class TCLinear(torch.nn.Linear):
    """``nn.Linear`` whose matmuls run on ``csrc/gemm_tcgen05.cu`` (bf16 in, fp32 accumulate, bf16 out)."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)


def swap_linear_modules(module: torch.nn.Module, min_features: int = 64, skip=()):
    """Replace every ``nn.Linear`` (in/out features ≥ ``min_features``, in_features % 8 == 0) by a ``TCLinear`` that
    shares its parameters.  Returns the number of layers swapped."""
    n = 0
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.Linear) and not isinstance(child, TCLinear) and name not in skip \
                and child.in_features % 8 == 0 and min(child.in_features, child.out_features) >= min_features:
            new = TCLinear(child.in_features, child.out_features, bias=child.bias is not None,
                           device=child.weight.device, dtype=child.weight.dtype)
            new.weight, new.bias = child.weight, child.bias
            setattr(module, name, new)
            n += 1
        else:
            n += swap_linear_modules(child, min_features, skip)
    return n
