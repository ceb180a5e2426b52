"""Server/client optimizers with the reference's update rules.

``AdamW`` (ref. ``utils/optimizers/adamW.py:30-87``: optional bias correction,
decoupled weight decay applied *after* the Adam step), ``LAMB``
(``lamb.py:69-134``: no debiasing, weight-norm clamp to 10, trust ratio),
``LarsSGDV1`` / ``LarsSGD`` (``lars.py:10-71`` / ``:74-128``) and a native
replacement for the external ``torchlars.LARS`` wrapper the reference uses for
``type: lars`` (``utils/utils.py:41-46``).

All of them are written multi-tensor style (``torch._foreach_*``): per-tensor
norms come out of ONE ``_foreach_norm`` launch and stay on the device, so a
step never synchronises with the host — the reference's versions call
``.norm()`` / ``== 0`` per tensor, i.e. one device→host sync per parameter.
The same formulas are implemented by the fused flat-arena CUDA kernels in
``msrflute_b200/csrc/server_update.cu``; these classes
are the CPU path and the numerical oracle for those kernels.
"""
from __future__ import annotations

import math

import torch
from torch.optim import Optimizer


def _grads_params(group):
    ps = [p for p in group["params"] if p.grad is not None]
    for p in ps:
        if p.grad.is_sparse:
            raise RuntimeError("sparse gradients are not supported")
    return ps, [p.grad for p in ps]


class AdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        if eps < 0.0:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                      correct_bias=correct_bias))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, gs = _grads_params(group)
            if not ps:
                continue
            b1, b2 = group["betas"]
            ms, vs = [], []
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            torch._foreach_mul_(ms, b1)
            torch._foreach_add_(ms, gs, alpha=1.0 - b1)
            torch._foreach_mul_(vs, b2)
            torch._foreach_addcmul_(vs, gs, gs, value=1.0 - b2)
            denom = torch._foreach_sqrt(vs)
            torch._foreach_add_(denom, group["eps"])
            t = self.state[ps[0]]["step"]
            step_size = group["lr"]
            if group["correct_bias"]:
                step_size = step_size * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
            torch._foreach_addcdiv_(ps, ms, denom, value=-step_size)
            if group["weight_decay"] > 0.0:
                torch._foreach_mul_(ps, 1.0 - group["lr"] * group["weight_decay"])
        return loss


class LAMB(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False):
        if lr < 0.0 or eps < 0.0:
            raise ValueError("Invalid lr/eps")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        self.adam = adam
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, gs = _grads_params(group)
            if not ps:
                continue
            b1, b2 = group["betas"]
            ms, vs = [], []
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            torch._foreach_mul_(ms, b1)
            torch._foreach_add_(ms, gs, alpha=1 - b1)
            torch._foreach_mul_(vs, b2)
            torch._foreach_addcmul_(vs, gs, gs, value=1 - b2)
            denom = torch._foreach_sqrt(vs)
            torch._foreach_add_(denom, group["eps"])
            adam_step = torch._foreach_div(ms, denom)
            if group["weight_decay"] != 0:
                torch._foreach_add_(adam_step, ps, alpha=group["weight_decay"])
            w_norm = torch.stack(torch._foreach_norm(ps)).clamp_(0, 10)
            a_norm = torch.stack(torch._foreach_norm(adam_step))
            trust = torch.where((w_norm == 0) | (a_norm == 0), torch.ones_like(w_norm), w_norm / a_norm)
            for i, p in enumerate(ps):          # diagnostics stay device tensors (no sync)
                st = self.state[p]
                st["weight_norm"], st["adam_norm"], st["trust_ratio"] = w_norm[i], a_norm[i], trust[i]
            if self.adam:
                trust = torch.ones_like(trust)
            scale = (-group["lr"] * trust).unbind(0)
            torch._foreach_mul_(adam_step, scale)
            torch._foreach_add_(ps, adam_step)
        return loss


class LarsSGDV1(torch.optim.SGD):
    """LARS (arXiv:1708.03888): lr_layer = min(5, lr·0.001·‖w‖/(‖g‖+wd‖w‖))."""

    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        super().__init__(params, lr=lr, momentum=momentum, dampening=dampening,
                         weight_decay=weight_decay, nesterov=nesterov)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, gs = _grads_params(group)
            if not ps:
                continue
            wd, mom, nesterov = group["weight_decay"], group["momentum"], group["nesterov"]
            p_n = torch.stack(torch._foreach_norm(ps))
            g_n = torch.stack(torch._foreach_norm(gs))
            if wd != 0:
                g_n = g_n + wd * p_n
                torch._foreach_add_(gs, ps, alpha=wd)
            lrs = torch.clamp(0.001 * p_n / g_n * group["lr"], max=5.0).unbind(0)
            upd = []
            for p, g, lr in zip(ps, gs, lrs):
                d_p = g
                if mom != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        buf = st["momentum_buffer"] = torch.clone(g).detach()
                    else:
                        buf = st["momentum_buffer"]
                        buf.mul_(mom).add_(g * lr)
                    d_p = g.add(buf, alpha=mom) if nesterov else buf
                upd.append(d_p)
            torch._foreach_sub_(ps, upd)
        return loss


class LarsSGD(torch.optim.SGD):
    """LARS as in arXiv:1904.00962 Alg. 1: lr_layer = clamp(lr·‖w‖/(‖u‖+1e-8), 0, 10)."""

    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        super().__init__(params, lr=lr, momentum=momentum, dampening=dampening,
                         weight_decay=weight_decay, nesterov=nesterov)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, gs = _grads_params(group)
            if not ps:
                continue
            mom, nesterov = group["momentum"], group["nesterov"]
            # NB: the reference computes ``d_p.add(p, alpha=wd)`` out of place and
            # discards the result (lars.py:104), i.e. weight decay is a no-op there.
            upd = []
            for p, g in zip(ps, gs):
                d_p = g
                if mom != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        buf = st["momentum_buffer"] = torch.clone(g).detach()
                    else:
                        buf = st["momentum_buffer"]
                        buf.mul_(mom).add_(g, alpha=1 - mom)
                    d_p = g.add(buf, alpha=mom) if nesterov else buf
                upd.append(d_p)
            p_n = torch.stack(torch._foreach_norm(ps))
            u_n = torch.stack(torch._foreach_norm(upd))
            lrs = (group["lr"] * p_n / (u_n + 1e-8)).clamp_(0, 10)
            scaled = torch._foreach_mul(upd, lrs.unbind(0))
            torch._foreach_sub_(ps, scaled)
        return loss


class LARSWrapper(Optimizer):
    """Layer-wise adaptive rate scaling around a base optimizer.

    Stand-in for ``torchlars.LARS(optimizer, eps, trust_coef)``: before the base
    step every gradient is rescaled by
    ``trust_coef·‖w‖ / (‖g‖ + wd·‖w‖ + eps)`` (1 when either norm is 0) and the
    weight-decay term is folded into the gradient.
    """

    def __init__(self, optimizer, eps=1e-8, trust_coef=0.001):
        self.optim = optimizer
        self.eps, self.trust_coef = eps, trust_coef
        self.param_groups = optimizer.param_groups
        self.state = optimizer.state
        self.defaults = optimizer.defaults

    def state_dict(self):
        return self.optim.state_dict()

    def load_state_dict(self, sd):
        self.optim.load_state_dict(sd)

    def zero_grad(self, set_to_none=True):
        self.optim.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        saved_wd = []
        for group in self.param_groups:
            wd = group.get("weight_decay", 0)
            saved_wd.append(wd)
            ps, gs = _grads_params(group)
            if not ps:
                continue
            p_n = torch.stack(torch._foreach_norm(ps))
            g_n = torch.stack(torch._foreach_norm(gs))
            ratio = self.trust_coef * p_n / (g_n + wd * p_n + self.eps)
            ratio = torch.where((p_n > 0) & (g_n > 0), ratio, torch.ones_like(ratio))
            if wd != 0:
                torch._foreach_add_(gs, ps, alpha=wd)
            torch._foreach_mul_(gs, ratio.unbind(0))
            group["weight_decay"] = 0
        loss = self.optim.step(closure)
        for group, wd in zip(self.param_groups, saved_wd):
            group["weight_decay"] = wd
        return loss


__all__ = ["AdamW", "LAMB", "LarsSGD", "LarsSGDV1", "LARSWrapper"]
