"""Differential privacy for FL rounds (ref. ``extensions/privacy/__init__.py``).

* ``apply_local_dp``  (ref :154-201) — per-client: clip (eps<0) or normalise-to-``max_grad`` + Gaussian
  noise with σ = √(2 ln(1.25/δ))·S/ε on [gradient ‖ weight], weight clamp and un-scaling.
* ``apply_global_dp`` (ref :128-151) — server: ``g += N(0,1)·global_sigma·max_grad/num_clients``.
* ``update_privacy_accountant`` (ref :204-260) — RDP accounting of the sampled Gaussian mechanism.
* PrivUnit2 / scalar-DP / Laplace mechanisms (ref :51-101).

The gradient math runs on the model's flat gradient arena: no ``torch.cat`` of all layers, no
re-scatter (the reference materialises three flat copies per client), and the norm is a device scalar.
On the fused device path the same formulas are applied by ``csrc/server_update.cu`` (global noise inside the
reduce/update kernel, Philox counter = element index, so the noise is independent of GPU count and tiling).
"""
import json
import logging
import math

import numpy as np
import torch as T
from scipy.special import betainc, betaln

from ...ops import misc_ops
from ...utils import print_rank
from . import analysis as privacy_analysis
from . import rng
from .analysis import RDPIncrementalAccountant

ORDERS = [1.25, 1.5, 1.75, 2., 2.25, 2.5, 3., 3.5, 4., 4.5] + list(range(5, 64)) + [128, 256, 512]
_ACCOUNTANT = RDPIncrementalAccountant(ORDERS)


def compute_LDP_noise_std(eps, max_sensitivity, delta):
    return np.sqrt(2 * np.log(1.25 / delta)) * max_sensitivity / eps


# ----------------------------------------------------------------- PrivUnit2
def _beta2betainc_ratio(a, x):
    return 1 / betainc(a, a, x)


def _log_m1(d, alpha, gamma):
    return alpha * np.log(1 - gamma ** 2) - (d - 2) * np.log(2) - np.log(d - 1)


def _log_m2(p, tau, alpha):
    r = _beta2betainc_ratio(alpha, tau)
    return np.log(p / (r - 1) - (1 - p)) + np.log(r) - betaln(alpha, alpha)


def _efficient_m(d, gamma, p):
    alpha, tau = (d - 1) / 2, (1 + gamma) / 2
    return np.exp(_log_m1(d, alpha, gamma) + _log_m2(p, tau, alpha))


def privacy_parameters(eps0, eps, d):
    e0, e = np.exp(eps0), np.exp(eps)
    p0 = 1 if e0 == np.inf else e0 / (1 + e0)
    base = np.sqrt(np.pi / (2 * (d - 1)))
    gamma = base if e == np.inf else ((e - 1) / (e + 1)) * base
    return p0, gamma


def private_unit2(grad, gamma, prob):
    """PrivUnit2 (Bhowmick et al.): sample a unit vector from the spherical cap {v: <v,g> ≥ γ} w.p. ``prob``
    (else from its complement) and debias by 1/m."""
    np.testing.assert_almost_equal(grad.norm().cpu().item(), 1, decimal=5)
    assert prob >= 0.5 and 0 <= gamma <= 1
    want_cap = bool(rng.rand(()) < prob)
    while True:
        V = rng.randn(grad.shape, device=grad.device)
        V = V / V.norm()
        if bool(T.dot(V, grad) >= gamma) == want_cap:
            break
    return V / _efficient_m(grad.shape[0], gamma, prob)


def add_private_unit2_noise(eps, grad):
    p, gamma = privacy_parameters(0.01 * eps, 0.99 * eps, grad.shape[0])
    return private_unit2(grad, gamma, p)


def add_gaussian_noise(grad, eps, max_grad, delta):
    sigma = compute_LDP_noise_std(eps, max_grad, delta)
    return grad + sigma * rng.randn(grad.shape, device=grad.device), sigma


def scalar_DP(r, eps, k, r_max):
    """ε-LDP release of a bounded scalar by randomised rounding to k+1 levels + k-ary randomised response."""
    r = np.minimum(r, r_max)
    val = k * r / r_max
    lo, hi = math.floor(val), math.ceil(val)
    J = lo if rng.rand(()) < (hi - val) else hi
    e = np.exp(eps)
    if rng.rand(()) >= e / (e + k):
        while True:
            J_ = int(T.randint(0, k + 1, (), generator=rng.dp_generator()).item())
            if J_ != J:
                J = J_
                break
    a = ((e + k) / (e - 1)) * (r_max / k)
    b = (k * (k + 1)) / (2 * (e + k))
    return a * (J - b)


def laplace_noise(max_sens, eps, vocab_size):
    u = rng.rand((int(vocab_size),)).double().numpy() - 0.5          # inverse-CDF sampling from the DP generator
    return -(max_sens / eps) * np.sign(u) * np.log1p(-2.0 * np.abs(u))


# ------------------------------------------------------------ flat grad access
def unroll_network(named_params, select_grad=False):
    """Flatten params/grads to one vector + index map (kept for API parity; prefer the arena)."""
    ids, flats, cur = {}, [], 0
    for n, p in named_params:
        d = (p.grad if select_grad else p.data).reshape(-1)
        flats.append(d)
        ids[n] = (cur, cur + d.shape[0])
        cur += d.shape[0]
    return T.cat(flats), ids


def update_network(named_params, params_ids, flat_params, apply_to_grad=False):
    for n, p in named_params:
        s, e = params_ids[n]
        (p.grad if apply_to_grad else p.data).copy_(flat_params[s:e].view_as(p))


def _flat_grad(model):
    """(flat_view, writer) — zero-copy when the model is arena-backed."""
    from ...core.strategies.utils import grad_arena
    ga = grad_arena(model)
    if ga is not None:
        return ga.flat, None
    flat, ids = unroll_network(model.named_parameters(), select_grad=True)
    return flat, (lambda f: update_network(model.named_parameters(), ids, f, apply_to_grad=True))


# -------------------------------------------------------------------- global
def apply_global_dp(config, model, num_clients_curr_iter, select_grad=True, metric_logger=None):
    dp = config.get("dp_config", None)
    if dp is None or not dp.get("enable_global_dp", False):
        return None
    assert dp["enable_local_dp"], "global DP requires enable_local_dp (client-side clipping)"
    if select_grad:
        flat, writer = _flat_grad(model)
    else:
        flat, ids = unroll_network(model.named_parameters(), select_grad=False)
        writer = lambda f: update_network(model.named_parameters(), ids, f, apply_to_grad=False)
    noise_scale = dp["global_sigma"] * dp["max_grad"] / num_clients_curr_iter
    grad_norm = flat.norm()
    if writer is None:
        flat.add_(rng.randn_like(flat), alpha=noise_scale)
    else:
        writer(flat + rng.randn_like(flat) * noise_scale)
    if metric_logger is not None:
        metric_logger("Gradient Norm", grad_norm.item())
    return noise_scale


# --------------------------------------------------------------------- local
def apply_local_dp(trainer, weight, dp_config, add_weight_noise):
    """Client-side DP on the pseudo-gradient held in ``trainer.model``'s grads; returns the (noisy) weight."""
    flat, writer = _flat_grad(trainer.model)
    grad_norm = flat.norm()
    max_grad = dp_config["max_grad"]
    fused = writer is None and flat.is_cuda and flat.is_contiguous() and flat.dtype == T.float32
    if dp_config["eps"] < 0:
        if fused:                                            # clip only: one reduction + one scale kernel
            misc_ops.local_dp_(flat, max_grad, 0.0, True)
            return weight
        coef = T.clamp(max_grad / grad_norm, max=1.0)        # clip only, stays on device
        out = flat.mul_(coef) if writer is None else flat * coef
        if writer is not None:
            writer(out)
        return weight
    eps, delta = dp_config["eps"], dp_config.get("delta", 1e-7)
    scaler = dp_config.get("weight_scaler", 1)
    weight_in = weight
    w_scaled = min(dp_config["max_weight"], scaler * weight)
    sens = np.sqrt(max_grad ** 2 + (dp_config["max_weight"] ** 2 if add_weight_noise else 0.0))
    sigma = compute_LDP_noise_std(eps, sens, delta)
    scale = max_grad / grad_norm
    if fused:                                                # scale to C and add Philox Gaussian noise in one pass
        misc_ops.local_dp_(flat, max_grad, float(sigma), False, seed=rng.dp_seed(stream=1))
    elif writer is None:
        flat.mul_(scale).add_(rng.randn_like(flat), alpha=float(sigma))
    else:
        writer(flat * scale + float(sigma) * rng.randn_like(flat))
    noisy_w = w_scaled + float(sigma) * rng.randn(()).item()
    weight = min(max(noisy_w, dp_config["min_weight"]), dp_config["max_weight"]) / scaler
    if not add_weight_noise:
        weight = weight_in
    print_rank("weight is {} and noisy weight is {}".format(weight_in, weight), loglevel=logging.DEBUG)
    return weight


# ---------------------------------------------------------------- accounting
def update_privacy_accountant(config, num_clients, curr_iter, num_clients_curr_iter, metric_logger=None):
    dp = config.get("dp_config", None)
    if dp is None or not (dp.get("enable_global_dp", False) or dp.get("enable_local_dp", False)):
        return None
    K, B, n, Tn = 1, num_clients_curr_iter, num_clients, curr_iter + 1
    delta = dp.get("delta", min(1e-7, 1. / (n * math.log(n))) if n > 1 else 1e-7)
    if dp.get("global_sigma", None) is None:
        sens = np.sqrt(dp["max_grad"] ** 2 + dp["max_weight"] ** 2)
        noise_scale = compute_LDP_noise_std(dp["eps"], sens, delta)
        global_sigma = noise_scale * np.sqrt(B) / sens
    else:
        global_sigma = dp["global_sigma"]
        noise_scale = global_sigma * dp["max_grad"] / B
    try:
        mu = K * B / n * math.sqrt(Tn * math.exp((1. / global_sigma) ** 2 - 1))
    except (OverflowError, ValueError, ZeroDivisionError):
        mu = -1
    q = B / n
    rdp_eps, opt_order = _ACCOUNTANT.epsilon(q, global_sigma, Tn, delta)
    props = {"dp_global_K": K, "dp_global_B": B, "dp_global_n": n, "dp_global_T": Tn,
             "dp_sigma": float(global_sigma), "dp_global_mu": mu, "dp_epsilon_rdp": rdp_eps,
             "dp_opt_order": opt_order, "dp_delta": delta, "dp_noise_scale": float(noise_scale)}
    print_rank(f"DP accounting: {json.dumps(props)}", logging.DEBUG)
    if metric_logger is None:
        from ...utils.metrics_sink import get_run
        metric_logger = get_run().log
    for k, v in props.items():
        metric_logger(k, v)
    return rdp_eps
