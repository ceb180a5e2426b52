"""Python faces of ``csrc/misc_kernels.cu`` with a PyTorch fp32 reference of each op — the CPU
path and the test oracle."""
import torch

from . import _ext


# ------------------------------------------------------------------------------------------------ K15 local DP
def local_dp_(flat: torch.Tensor, max_grad: float, sigma: float, clip_only: bool, seed: int = 0) -> torch.Tensor:
    """In place ``g ← g·s + σ·N(0,1)`` with ``s = min(1, C/‖g‖)`` (``clip_only``) or ``C/‖g‖`` (the reference's
    Gaussian mechanism, ``extensions/privacy/__init__.py:154-201``).  Returns ``‖g‖`` before scaling (device scalar).
    CUDA: one reduction + one fused scale/noise kernel, Philox noise keyed by element index."""
    if _ext.use_cuda_kernels(flat) and flat.is_contiguous() and flat.dtype == torch.float32:
        norm = _ext.load().local_dp(flat.view(-1), float(max_grad), float(sigma), bool(clip_only), int(seed))
        _ext.count_launch(2)
        return norm.reshape(())
    return local_dp_reference_(flat, max_grad, sigma, clip_only, seed)


def local_dp_reference_(flat, max_grad, sigma, clip_only, seed=0):
    norm = flat.norm()
    scale = max_grad / norm
    if clip_only:
        scale = torch.clamp(scale, max=1.0)
    flat.mul_(scale)
    if sigma != 0.0:
        gen = torch.Generator(device=flat.device)
        gen.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
        flat.add_(torch.randn(flat.shape, generator=gen, device=flat.device, dtype=flat.dtype), alpha=float(sigma))
    return norm


# ------------------------------------------------------------------------------------------------ K7 softmax-CE
class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        ext = _ext.load()
        loss, dx = ext.softmax_ce(logits.contiguous(), target.contiguous(), 1.0, int(ignore_index), True)
        _ext.count_launch(1)
        ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dx,) = ctx.saved_tensors
        return dx * dloss.unsqueeze(1), None, None


def softmax_cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Per-row cross entropy ``[rows]`` of fp32 ``logits [rows, C]``.  CUDA: the forward kernel also produces
    ``softmax − onehot`` so the backward is a single scale (one launch instead of log_softmax + nll fwd/bwd)."""
    if _ext.use_cuda_kernels(logits) and logits.dtype == torch.float32 and logits.dim() == 2:
        return _SoftmaxCE.apply(logits, target.long(), ignore_index)
    return torch.nn.functional.cross_entropy(logits.float(), target.long(), reduction="none", ignore_index=ignore_index)


# ------------------------------------------------------------------------------------------------ K19 cosine
def cosine_stats(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``[<a,b>, ‖a‖², ‖b‖²]`` of two flat fp32 tensors in one pass."""
    if _ext.use_cuda_kernels(a) and a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype == torch.float32:
        _ext.count_launch(1)
        return _ext.load().cosine_stats(a.view(-1), b.view(-1))
    a, b = a.reshape(-1).float(), b.reshape(-1).float()
    return torch.stack([torch.dot(a, b), torch.dot(a, a), torch.dot(b, b)])


def cosine(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    s = cosine_stats(a, b)
    den = torch.sqrt(s[1] * s[2])
    return torch.where(den > 0, s[0] / den.clamp(min=1e-30), torch.zeros_like(den))


# ------------------------------------------------------------------------------------------------ K25 personalization
def alpha_dot(wp: torch.Tensor, wg: torch.Tensor, gp: torch.Tensor, gg: torch.Tensor, alpha: float) -> torch.Tensor:
    """``sum((wp - wg) * (alpha * gp + (1 - alpha) * gg))`` over four flat fp32 vectors (personal / global weights and
    gradients) as one 1-element float64 tensor.  CUDA: one pass over the four arenas (csrc/gather_kernels.cu)."""
    if _ext.use_cuda_kernels(wp) and all(t.is_contiguous() and t.dtype == torch.float32 and t.data_ptr() % 16 == 0
                                         for t in (wp, wg, gp, gg)):
        _ext.count_launch(1)
        return _ext.load().alpha_dot(wp.view(-1), wg.view(-1), gp.view(-1), gg.view(-1), float(alpha))
    wp, wg, gp, gg = (t.reshape(-1).double() for t in (wp, wg, gp, gg))
    return torch.dot(wp - wg, alpha * gp + (1.0 - alpha) * gg).reshape(1)


# ------------------------------------------------------------------------------------------------ K8 max-pool
class _MaxPool2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        ext = _ext.load()
        x = x.contiguous()
        y, arg = ext.max_pool2d_fwd(x, int(k), int(stride), int(pad))
        _ext.count_launch(1)
        ctx.save_for_backward(arg)
        ctx.cfg = (x.shape[-2], x.shape[-1], int(k), int(stride), int(pad))
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        H, W, k, stride, pad = ctx.cfg
        dx = _ext.load().max_pool2d_bwd(dy.contiguous(), arg, H, W, k, stride, pad)
        _ext.count_launch(1)
        return dx, None, None, None


def max_pool2d(x: torch.Tensor, kernel_size: int, stride: int, padding: int = 0) -> torch.Tensor:
    """NCHW fp32 max-pool; CUDA: one forward kernel that also stores a 1-byte window argmax, one backward kernel."""
    if _ext.use_cuda_kernels(x) and x.dtype == torch.float32 and x.dim() >= 3 and hasattr(_ext.load(), "max_pool2d_fwd"):
        return _MaxPool2d.apply(x, kernel_size, stride, padding)
    return torch.nn.functional.max_pool2d(x, kernel_size, stride, padding)
