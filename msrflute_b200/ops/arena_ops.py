"""Flat-arena operations: the non-GEMM hot paths of a federated round.

Every function here has two implementations with identical semantics:

* a hand-written sm_100a CUDA kernel (``csrc/arena_kernels.cu``,
  ``csrc/server_update.cu``) used whenever the tensors live on a GPU, and
* the PyTorch reference right below it — the CPU/gloo path and the numerical
  oracle the GPU tests compare against.

Semantics are taken from the reference call sites:

``fused_client_step``  ``clip_grad_norm_`` (``core/trainer.py:383-384``) +
                       ``estimate_sufficient_stats`` (``:263-312``; the reference
                       copies every gradient to the host each mini-batch) +
                       ``optimizer.step()`` (``:391``) + ``zero_grad`` (``:368``).
``accumulate_pseudo_grad``  ``p.grad = w_global − w_local`` (``core/client.py:380-383``),
                       ``weight·grad`` (``core/strategies/fedavg.py:80``) and the server-side
                       ``p.grad += client_grad`` (``core/strategies/utils.py:21-33``).
``server_update``      ``p.grad /= weight_sum`` (``fedavg.py:146-147``), global DP noise
                       (``extensions/privacy/__init__.py:128-151``), server clip + optimizer step
                       (``core/trainer.py:127-137``) and the next round's weight broadcast
                       (``core/federated.py:330-334``) — over peer memory when ``peers`` are given.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import _ext

# hyper-parameter row layout (device-resident so CUDA graphs survive lr changes)
H_LR, H_MAXNORM, H_WD, H_MOM = 0, 1, 2, 3
# stats row layout
S_SUM, S_SUMSQ, S_COUNT, S_LASTNORM = 0, 1, 2, 3


def make_hyper(rows: int, device, lr=0.0, max_norm=0.0, weight_decay=0.0, momentum=0.0) -> torch.Tensor:
    h = torch.zeros(rows, 4, dtype=torch.float32, device=device)
    h[:, H_LR], h[:, H_MAXNORM], h[:, H_WD], h[:, H_MOM] = lr, max_norm or 0.0, weight_decay, momentum
    return h


def _2d(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return t if t.dim() == 2 else t.view(1, -1)


# --------------------------------------------------------------- client step
def fused_client_adamw(w, g, m, v, step, hyper, stats, *, n_logical: int, betas=(0.9, 0.999), eps: float = 1e-6,
                       correct_bias: bool = True, zero_grad: bool = True):
    """clip → sufficient stats → AdamW (HF / reference semantics: ``eps`` outside the bias correction, decoupled weight
    decay applied after the update) → zero grad, per arena row.  ``m, v``: ``[S, P]`` moment arenas, ``step``: ``[S]``
    int32 device tensor holding the 1-based step count of this update, ``hyper``: ``[S, 4]`` (lr, max_norm, wd, -)."""
    w2, g2, m2, v2 = _2d(w), _2d(g), _2d(m), _2d(v)
    b1, b2 = float(betas[0]), float(betas[1])
    if _ext.use_cuda_kernels(w2, g2, hyper, stats):
        _ext.load().fused_client_adamw(w2, g2, m2, v2, step, hyper, stats, int(n_logical), b1, b2, float(eps),
                                       bool(correct_bias), bool(zero_grad))
        _ext.count_launch(2)
        return
    sumsq = (g2 * g2).sum(dim=1)
    ssum = g2.sum(dim=1)
    norm = sumsq.sqrt()
    max_norm = hyper[:, H_MAXNORM]
    coef = torch.where(max_norm > 0, (max_norm / (norm + 1e-6)).clamp(max=1.0), torch.ones_like(norm))
    stats[:, S_SUM] += coef * ssum
    stats[:, S_SUMSQ] += coef * coef * sumsq
    stats[:, S_COUNT] += float(n_logical)
    stats[:, S_LASTNORM] = norm
    gg = g2 * coef[:, None]
    m2.mul_(b1).add_(gg, alpha=1.0 - b1)
    v2.mul_(b2).addcmul_(gg, gg, value=1.0 - b2)
    lr, wd = hyper[:, H_LR, None], hyper[:, H_WD, None]
    t = step.to(torch.float32).view(-1, 1)
    step_size = lr * (torch.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)) if correct_bias else lr
    w2.sub_(step_size * m2 / (v2.sqrt() + eps))
    w2.mul_(torch.where(wd > 0, 1.0 - lr * wd, torch.ones_like(wd)))
    if zero_grad:
        _2d(g).zero_()


def fused_client_step(w, g, hyper, stats, mom=None, *, n_logical: int, nesterov: bool = False,
                      dampening: float = 0.0, zero_grad: bool = True, first_step=None, prox_ref=None, prox_mult=None,
                      prox_loss=None):
    """clip → sufficient stats → SGD(momentum, weight-decay) → zero grad, per arena row.

    ``w, g, mom``: ``[S, P]`` (or ``[P]``) fp32.  ``hyper``: ``[S, 4]`` (lr, max_norm, wd, momentum), a
    ``max_norm <= 0`` disables clipping.  ``stats``: ``[S, 4]`` accumulators (Σg, Σg², n, last ‖g‖), the
    sums are over the *clipped* gradient like the reference's.  ``first_step``: optional ``[S]`` int32 device
    flags — rows whose momentum buffer must be initialised with the gradient (torch.optim.SGD semantics).
    ``prox_ref`` / ``prox_mult`` (``[P]`` each): FedProx — the gradient of the proximal term,
    ``prox_mult * (w - prox_ref)``, joins ``g`` before clipping and statistics (``prox_mult`` = mu times
    the reference's per-tensor multiplicity, 0 on padding).
    """
    w2, g2, m2 = _2d(w), _2d(g), _2d(mom)
    if _ext.use_cuda_kernels(w2, g2, hyper, stats):
        _ext.load().fused_client_step(w2, g2, hyper, stats, m2, first_step, int(n_logical), bool(nesterov),
                                      float(dampening), bool(zero_grad), prox_ref, prox_mult, prox_loss)
        _ext.count_launch(2)
        return
    if prox_ref is not None and prox_mult is not None:
        diff = w2 - prox_ref.view(1, -1)
        if prox_loss is not None:
            prox_loss.add_(0.5 * (prox_mult.view(1, -1) * diff * diff).sum(dim=1))
        g2 = g2 + prox_mult.view(1, -1) * diff
        g_out = _2d(g)
    else:
        g_out = g2
    sumsq = (g2 * g2).sum(dim=1)
    ssum = g2.sum(dim=1)
    norm = sumsq.sqrt()
    max_norm = hyper[:, H_MAXNORM]
    coef = torch.where(max_norm > 0, (max_norm / (norm + 1e-6)).clamp(max=1.0), torch.ones_like(norm))
    stats[:, S_SUM] += coef * ssum
    stats[:, S_SUMSQ] += coef * coef * sumsq
    stats[:, S_COUNT] += float(n_logical)
    stats[:, S_LASTNORM] = norm
    d = g2 * coef[:, None] + hyper[:, H_WD, None] * w2
    if m2 is not None:
        mu = hyper[:, H_MOM, None]
        if first_step is not None:
            fs = first_step.view(-1, 1).bool()
            new_m = torch.where(fs, d, mu * m2 + (1.0 - dampening) * d)
        else:
            new_m = mu * m2 + (1.0 - dampening) * d
        m2.copy_(new_m)
        d = d + mu * new_m if nesterov else new_m
    w2.sub_(hyper[:, H_LR, None] * d)
    if zero_grad:
        g_out.zero_()


# ---------------------------------------------------------------------------------------------- slot-layout gather
def slot_gather_bcast(W, wg_slot, wg, index_map):
    """Model distribution into the slot arenas: ``wg_slot[j] = wg[map[j]]`` (0 on padding), ``W[s] = wg_slot`` ∀ s."""
    if _ext.use_cuda_kernels(W, wg_slot, wg, index_map):
        _ext.load().slot_gather_bcast(W, wg_slot, wg, index_map)
        _ext.count_launch(1)
        return
    m = index_map.long()
    row = torch.where(m >= 0, wg[m.clamp(min=0)], torch.zeros((), dtype=wg.dtype, device=wg.device))
    wg_slot.copy_(row)
    W.copy_(row.view(1, -1).expand_as(W))


def slot_pg_sqnorm(W, wg_slot, out):
    """``out[s] = ||wg_slot - W[s]||^2`` (pseudo-gradient norms for local-DP clipping / normalisation)."""
    out.zero_()
    if _ext.use_cuda_kernels(W, wg_slot, out):
        _ext.load().slot_pg_sqnorm(W, wg_slot, out)
        _ext.count_launch(1)
        return out
    out.copy_(((wg_slot.view(1, -1) - W) ** 2).sum(dim=1))
    return out


def slot_gather_fused(acc_slot, W, wg_slot, coef, sig=None, seed=None):
    """``acc_slot += Σ_s coef[s]·(wg_slot − W[s]) + Σ_s sig[s]·N(seed[s])`` — weighting, local-DP scaling and Gaussian
    noise of every client in one pass (all coefficients are device tensors)."""
    if _ext.use_cuda_kernels(acc_slot, W, wg_slot, coef):
        _ext.load().slot_gather_fused(acc_slot, W, wg_slot, coef, sig, seed)
        _ext.count_launch(1)
        return
    acc_slot.add_((coef.view(-1, 1) * (wg_slot.view(1, -1) - W)).sum(dim=0))
    if sig is not None and seed is not None:
        for s in range(W.shape[0]):
            if float(sig[s]) != 0.0:
                acc_slot.add_(_philox_like_noise(acc_slot.numel(), int(seed[s]), acc_slot.device), alpha=float(sig[s]))


def slot_scatter_acc(acc, acc_slot, index_map):
    """``acc[map[j]] += acc_slot[j]`` then ``acc_slot = 0`` (the one pass through the permutation per round)."""
    if _ext.use_cuda_kernels(acc, acc_slot, index_map):
        _ext.load().slot_scatter_acc(acc, acc_slot, index_map)
        _ext.count_launch(1)
        return
    m = index_map.long()
    live = m >= 0
    acc.index_add_(0, m[live], acc_slot[live])
    acc_slot.zero_()


def dead_coord_noise(acc, dead_idx, sig2_sum, seed):
    """Local-DP noise for coordinates that no slot stores (elided dead filter taps): ``acc[idx] += sqrt(Σ sig²)·N``."""
    if dead_idx is None or dead_idx.numel() == 0:
        return
    if _ext.use_cuda_kernels(acc, dead_idx, sig2_sum):
        _ext.load().dead_coord_noise(acc, dead_idx, sig2_sum.reshape(1).float(), int(seed))
        _ext.count_launch(1)
        return
    noise = _philox_like_noise(dead_idx.numel(), seed, acc.device)
    acc.index_add_(0, dead_idx.long(), noise * float(sig2_sum.clamp(min=0).sqrt()))


def slot_quant_stats(W, wg_row, segs, q: float, bits: int):
    """Per (slot, tensor) quantization statistics of the pseudo-gradient ``wg_row − W[s]``: ``[S, T, 4]`` =
    (min, bin width, |g|-quantile threshold, max).  ``segs``: int64 ``[T, 3]`` (offset, stored elements, elements of the
    full tensor — the difference are structurally-zero entries the slot layout does not store).  CUDA: radix select
    (``csrc/quant_gather.cu``); the PyTorch path below (``torch.quantile`` on the full tensor) is the oracle."""
    if _ext.use_cuda_kernels(W, wg_row, segs):
        out = _ext.load().slot_quant_stats(W, wg_row, segs, float(q), int(bits))
        _ext.count_launch(9)
        return out
    S, T = W.shape[0], segs.shape[0]
    out = torch.zeros(S, T, 4, dtype=torch.float32, device=W.device)
    for s in range(S):
        for t, (off, n, ntot) in enumerate(segs.tolist()):
            d = (wg_row[off:off + n] - W[s, off:off + n]).float()
            full = torch.cat([d, d.new_zeros(ntot - n)]) if ntot > n else d
            lo, hi = full.min(), full.max()
            a = full.abs()
            if a.numel() > 2 ** 24:
                pos = q * (a.numel() - 1)
                srt = a.sort().values
                k0 = int(pos)
                thr = srt[k0] + (pos - k0) * (srt[min(k0 + 1, a.numel() - 1)] - srt[k0])
            else:
                thr = torch.quantile(a, q)
            out[s, t] = torch.stack([lo, (hi - lo) / (2 ** bits - 1), thr, hi])
    return out


def slot_quant_gather(acc_row, W, wg_row, coef, params, seg_of_blk, bits: int):
    """``acc_row[j] += Σ_s coef[s]·Q_{s,t(j)}(wg_row[j] − W[s, j])`` — binning to ``2**bits`` levels on ``[lo, hi]``,
    zeroing below the threshold and the aggregation weight in one pass (ref. ``extensions/quantization/quant.py:53-100``
    applied to every client's payload)."""
    if _ext.use_cuda_kernels(acc_row, W, wg_row, coef, params, seg_of_blk):
        _ext.load().slot_quant_gather(acc_row, W, wg_row, coef, params, seg_of_blk, int(bits))
        _ext.count_launch(1)
        return
    L = 2 ** bits
    seg = seg_of_blk.long().repeat_interleave(32)
    live = seg >= 0
    segc = seg.clamp(min=0)
    for s in range(W.shape[0]):
        c = float(coef[s])
        if c == 0.0:
            continue
        d = wg_row - W[s]
        lo, width, thr = params[s, segc, 0], params[s, segc, 1], params[s, segc, 2]
        safe = torch.where(width > 0, width, torch.ones_like(width))
        idx = torch.ceil((d - lo) / safe - 0.5).clamp(0, L - 1)
        binned = torch.where(width > 0, lo + idx * width, lo)
        qd = torch.where((d.abs() > thr) & live, binned, torch.zeros_like(d))
        acc_row.add_(qd, alpha=c)


def clip_and_stats(g, hyper, stats, *, n_logical: int):
    """Clip in place + accumulate stats, for client optimizers other than SGD (the step is then done by torch)."""
    g2 = _2d(g)
    if _ext.use_cuda_kernels(g2, hyper, stats):
        _ext.load().clip_and_stats(g2, hyper, stats, int(n_logical))
        _ext.count_launch(2)
        return
    sumsq = (g2 * g2).sum(dim=1)
    ssum = g2.sum(dim=1)
    norm = sumsq.sqrt()
    max_norm = hyper[:, H_MAXNORM]
    coef = torch.where(max_norm > 0, (max_norm / (norm + 1e-6)).clamp(max=1.0), torch.ones_like(norm))
    g2.mul_(coef[:, None])
    stats[:, S_SUM] += coef * ssum
    stats[:, S_SUMSQ] += coef * coef * sumsq
    stats[:, S_COUNT] += float(n_logical)
    stats[:, S_LASTNORM] = norm


def finalize_stats(stats: torch.Tensor):
    """(mean, mag, var, norm) from the accumulators, as ``Trainer.estimate_sufficient_stats`` defines them
    (``core/trainer.py:295-312``) — including its quirk that ``var = E[g²] − mag²`` is identically ~0."""
    n = stats[..., S_COUNT].clamp(min=1.0)
    mean = stats[..., S_SUM] / n
    ex2 = stats[..., S_SUMSQ] / n
    mag = ex2.sqrt()
    var = ex2 - mag * mag
    norm = stats[..., S_SUMSQ].sqrt()
    return mean, mag, var, norm


# ------------------------------------------------------------ pseudo-gradient
def pseudo_grad(w_global, w_local, out, weight: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None):
    """``out = (w_global − w_local)·weight`` for ONE client (per-client delivery mode); optional Σ/Σ² of the
    unweighted pseudo-gradient into ``stats[0:2]`` (``stats_on_smooth_grad``)."""
    if _ext.use_cuda_kernels(w_global, w_local, out):
        _ext.load().pseudo_grad(w_global, w_local, out, weight, stats)
        _ext.count_launch(1)
        return
    pg = w_global - w_local
    if stats is not None:
        stats[S_SUM] = pg.sum()
        stats[S_SUMSQ] = (pg * pg).sum()
    out.copy_(pg if weight is None else pg * weight)


def scatter_in(W: torch.Tensor, w_global: torch.Tensor, index_map: torch.Tensor):
    """``W[s, j] = w_global[index_map[j]]`` for every slot row (0 where ``index_map[j] < 0``) — the round's
    "broadcast" into compact slot arenas."""
    if _ext.use_cuda_kernels(W):
        _ext.load().slot_scatter_in(W, w_global, index_map)
        _ext.count_launch(1)
        return W
    idx = index_map.long().clamp(min=0)
    row = torch.where(index_map >= 0, w_global[idx], torch.zeros((), dtype=W.dtype, device=W.device))
    W.copy_(row.view(1, -1).expand_as(W))
    return W


def accumulate_pseudo_grad_mapped(acc, w_global, w_local, weights, active, index_map):
    """``acc[map[j]] += Σ_s weights[s]·active[s]·(w_global[map[j]] − w_local[s, j])`` (compact slot arenas)."""
    if _ext.use_cuda_kernels(acc):
        _ext.load().accumulate_pseudo_grad_mapped(acc, w_global, w_local, weights, active, index_map)
        _ext.count_launch(1)
        return acc
    keep = index_map >= 0
    idx = index_map[keep].long()
    w = weights * (active.to(weights.dtype) if active is not None else 1.0)
    delta = (w.view(-1, 1) * (w_global[idx].view(1, -1) - w_local[:, keep])).sum(dim=0)
    acc[idx] += delta
    return acc


def accumulate_pseudo_grad(acc, w_global, w_local, weights, active: Optional[torch.Tensor] = None):
    """``acc += Σ_s weights[s]·(w_global − w_local[s])`` over the active rows, one pass, weights on device."""
    wl = _2d(w_local)
    if _ext.use_cuda_kernels(acc, w_global, wl, weights):
        _ext.load().accumulate_pseudo_grad(acc, w_global, wl, weights, active)
        _ext.count_launch(1)
        return
    wts = weights.view(-1).to(acc.dtype)
    if active is not None:
        wts = wts * active.view(-1).to(acc.dtype)
    acc.add_(((w_global.view(1, -1) - wl) * wts[:, None]).sum(dim=0))


def axpy_(y, x, alpha: float):
    """``y += alpha·x`` on flat buffers (per-client server-side aggregation in p2p mode)."""
    y.add_(x, alpha=alpha)


# ------------------------------------------------------------- server update
OPT_SGD, OPT_ADAM, OPT_ADAMW, OPT_ADAMAX, OPT_LAMB, OPT_LARS = 0, 1, 2, 3, 4, 5
OPT_CODES = {"sgd": OPT_SGD, "adam": OPT_ADAM, "adamW": OPT_ADAMW, "adamax": OPT_ADAMAX, "lamb": OPT_LAMB,
             "LarsSGD": OPT_LARS}


class ServerOptState:
    """Optimizer state held in arenas (m, v) + scalar hyper-parameters for the fused server kernel."""

    def __init__(self, kind: str, numel: int, device, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 momentum=0.0, dampening=0.0, nesterov=False, correct_bias=True, amsgrad=False):
        if kind not in OPT_CODES:
            raise ValueError("optimizer {} has no fused server kernel".format(kind))
        if amsgrad:
            raise ValueError("amsgrad is not supported by the fused server kernel")
        self.kind, self.code = kind, OPT_CODES[kind]
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.momentum, self.dampening, self.nesterov = float(momentum), float(dampening), bool(nesterov)
        self.correct_bias = bool(correct_bias)
        self.step = 0
        need_m = kind != "sgd" or momentum != 0 or kind == "LarsSGD" and momentum != 0
        need_v = kind in ("adam", "adamW", "adamax", "lamb")
        self.m = torch.zeros(numel, dtype=torch.float32, device=device) if need_m else None
        self.v = torch.zeros(numel, dtype=torch.float32, device=device) if need_v else None


def _philox_like_noise(numel, seed, device):
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    return torch.randn(numel, generator=gen, device=device, dtype=torch.float32)


def server_update(w: torch.Tensor, accs: Sequence[torch.Tensor], weight_sum: torch.Tensor, opt: ServerOptState, *,
                  grad_out: Optional[torch.Tensor] = None, noise_scale: float = 0.0, seed: int = 0,
                  max_grad_norm: Optional[float] = None, segments: Optional[torch.Tensor] = None,
                  bcast: Optional[Sequence[torch.Tensor]] = None, zero_accs: bool = True,
                  stats_out: Optional[torch.Tensor] = None, peer_ptrs=None):
    """Fused gather → reduce → normalise → (DP noise) → (clip) → optimizer → broadcast.

    ``accs``: one weighted pseudo-gradient accumulator per rank (peer-mapped buffers on NVLink, or just the
    local one); ``weight_sum``: 0-dim device tensor Σ weights; ``bcast``: buffers (local + peers) that receive
    the new weights.  ``grad_out`` optionally keeps the aggregated gradient (cosine dumps, ``Gradient Norm``).
    ``stats_out[0]`` receives ‖g‖ before noise/clipping.
    """
    use_cuda = _ext.use_cuda_kernels(w, *accs)
    if use_cuda:
        _ext.load().server_update(w, list(accs), weight_sum, opt.m, opt.v, grad_out, segments,
                                  list(bcast) if bcast is not None else [], stats_out,
                                  opt.code, opt.step + 1, opt.lr, opt.betas[0], opt.betas[1], opt.eps,
                                  opt.weight_decay, opt.momentum, opt.dampening, opt.nesterov, opt.correct_bias,
                                  float(noise_scale), int(seed), float(max_grad_norm or 0.0), bool(zero_accs))
        opt.step += 1
        _ext.count_launch(3)
        return
    g = torch.zeros_like(w)
    for a in accs:
        g.add_(a.to(w.device))
    g.div_(weight_sum.to(w.device))
    if stats_out is not None:
        stats_out[0] = g.norm()
    if noise_scale > 0:
        g.add_(_philox_like_noise(g.numel(), seed, g.device), alpha=noise_scale)
    if max_grad_norm:
        n = g.norm()
        g.mul_((max_grad_norm / (n + 1e-6)).clamp(max=1.0))
    if grad_out is not None:
        grad_out.copy_(g)
    apply_optimizer_reference(w, g, opt, segments)
    if zero_accs:
        for a in accs:
            a.zero_()
    if bcast is not None:
        for b in bcast:
            if b.data_ptr() != w.data_ptr():
                b.copy_(w)


def apply_optimizer_reference(w, g, opt: ServerOptState, segments=None):
    """PyTorch reference of the fused optimizer epilogues (same math as ``utils/optimizers`` and torch.optim)."""
    opt.step += 1
    t, lr, (b1, b2), eps, wd = opt.step, opt.lr, opt.betas, opt.eps, opt.weight_decay
    if opt.kind == "sgd":
        d = g + wd * w if wd else g
        if opt.momentum:
            if t == 1:
                opt.m.copy_(d)
            else:
                opt.m.mul_(opt.momentum).add_(d, alpha=1 - opt.dampening)
            d = d + opt.momentum * opt.m if opt.nesterov else opt.m
        w.sub_(lr * d)
    elif opt.kind == "adam":
        d = g + wd * w if wd else g
        opt.m.mul_(b1).add_(d, alpha=1 - b1)
        opt.v.mul_(b2).addcmul_(d, d, value=1 - b2)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        w.addcdiv_(opt.m, (opt.v.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)
    elif opt.kind == "adamW":
        opt.m.mul_(b1).add_(g, alpha=1 - b1)
        opt.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        step = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t) if opt.correct_bias else lr
        w.addcdiv_(opt.m, opt.v.sqrt().add_(eps), value=-step)
        if wd > 0:
            w.mul_(1 - lr * wd)
    elif opt.kind == "adamax":
        d = g + wd * w if wd else g
        opt.m.mul_(b1).add_(d, alpha=1 - b1)
        torch.maximum(opt.v.mul_(b2), d.abs().add_(eps), out=opt.v)
        w.addcdiv_(opt.m, opt.v, value=-lr / (1 - b1 ** t))
    elif opt.kind in ("lamb", "LarsSGD"):
        if segments is None:
            raise ValueError("{} needs the per-tensor segment table".format(opt.kind))
        if opt.kind == "lamb":
            opt.m.mul_(b1).add_(g, alpha=1 - b1)
            opt.v.mul_(b2).addcmul_(g, g, value=1 - b2)
            u = opt.m / opt.v.sqrt().add(eps)
            if wd:
                u.add_(w, alpha=wd)
        else:
            u = g
            if opt.momentum:
                if t == 1:
                    opt.m.copy_(g)
                else:
                    opt.m.mul_(opt.momentum).add_(g, alpha=1 - opt.momentum)
                u = g + opt.momentum * opt.m if opt.nesterov else opt.m
        for off, n in segments.tolist():
            ws, us = w[off:off + n], u[off:off + n]
            wn, un = ws.norm(), us.norm()
            if opt.kind == "lamb":
                wn = wn.clamp(0, 10)
                trust = torch.where((wn == 0) | (un == 0), torch.ones_like(wn), wn / un)
                ws.sub_(lr * trust * us)
            else:
                ws.sub_((lr * wn / (un + 1e-8)).clamp(0, 10) * us)
    else:
        raise ValueError(opt.kind)
