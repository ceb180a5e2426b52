"""Rényi-DP accounting for the (Poisson-)sampled Gaussian mechanism.

Same interface as the reference's ``extensions/privacy/analysis.py`` (``compute_rdp`` :245,
``get_privacy_spent`` :275): RDP of order α for sampling rate q and noise multiplier σ, composed
over ``steps``; ε(δ) = min_α [ RDP(α) − log δ / (α − 1) ].

Written from the published algorithm (Mironov, Talwar, Zhang 2019, "Rényi Differential Privacy of
the Sampled Gaussian Mechanism", §3.3): integer α uses the binomial expansion

    A(α) = Σ_{k=0..α} C(α,k) (1−q)^{α−k} q^k exp((k²−k)/(2σ²))

and fractional α the two-sided series with erfc tails; everything in log space.
"""
from __future__ import annotations

import math
from typing import List, Tuple, Union

import numpy as np
from scipy import special


def _log_add(a: float, b: float) -> float:
    lo, hi = min(a, b), max(a, b)
    if lo == -np.inf:
        return hi
    return hi + math.log1p(math.exp(lo - hi))


def _log_sub(a: float, b: float) -> float:
    if a < b:
        raise ValueError("log_sub: result would be negative")
    if b == -np.inf:
        return a
    if a == b:
        return -np.inf
    try:
        return math.log(math.expm1(a - b)) + b
    except OverflowError:
        return a


def _log_erfc(x: float) -> float:
    return math.log(2) + special.log_ndtr(-x * 2 ** 0.5)


def _log_a_int(q: float, sigma: float, alpha: int) -> float:
    out = -np.inf
    for k in range(alpha + 1):
        term = (math.log(special.binom(alpha, k)) + k * math.log(q) + (alpha - k) * math.log(1 - q)
                + (k * k - k) / (2 * sigma ** 2))
        out = _log_add(out, term)
    return float(out)


def _log_a_frac(q: float, sigma: float, alpha: float) -> float:
    # A = E_{z~μ0}[(μ/μ0)^α], split at z0 where the two Gaussians' weighted densities cross
    log_a0, log_a1 = -np.inf, -np.inf
    z0 = sigma ** 2 * math.log(1 / q - 1) + 0.5
    i = 0
    while True:
        coef = special.binom(alpha, i)
        log_coef = math.log(abs(coef))
        j = alpha - i
        log_t0 = log_coef + i * math.log(q) + j * math.log(1 - q)
        log_t1 = log_coef + j * math.log(q) + i * math.log(1 - q)
        log_e0 = math.log(0.5) + _log_erfc((i - z0) / (math.sqrt(2) * sigma))
        log_e1 = math.log(0.5) + _log_erfc((z0 - j) / (math.sqrt(2) * sigma))
        log_s0 = log_t0 + (i * i - i) / (2 * sigma ** 2) + log_e0
        log_s1 = log_t1 + (j * j - j) / (2 * sigma ** 2) + log_e1
        if coef > 0:
            log_a0, log_a1 = _log_add(log_a0, log_s0), _log_add(log_a1, log_s1)
        else:
            log_a0, log_a1 = _log_sub(log_a0, log_s0), _log_sub(log_a1, log_s1)
        i += 1
        if max(log_s0, log_s1) < -30:
            break
    return _log_add(log_a0, log_a1)


def _compute_log_a(q: float, sigma: float, alpha: float) -> float:
    if float(alpha).is_integer():
        return _log_a_int(q, sigma, int(alpha))
    return _log_a_frac(q, sigma, alpha)


def _compute_rdp(q: float, sigma: float, alpha: float) -> float:
    if q == 0:
        return 0
    if sigma == 0:
        return np.inf
    if q == 1.0:
        return alpha / (2 * sigma ** 2)
    if np.isinf(alpha):
        return np.inf
    return _compute_log_a(q, sigma, alpha) / (alpha - 1)


def compute_rdp(q: float, noise_multiplier: float, steps: int, orders: Union[List[float], float]):
    if isinstance(orders, (float, int)):
        rdp = _compute_rdp(q, noise_multiplier, float(orders))
    else:
        rdp = np.array([_compute_rdp(q, noise_multiplier, o) for o in orders])
    return rdp * steps


def get_privacy_spent(orders, rdp, delta: float) -> Tuple[float, float]:
    orders_vec, rdp_vec = np.atleast_1d(orders).astype(float), np.atleast_1d(rdp).astype(float)
    if len(orders_vec) != len(rdp_vec):
        raise ValueError("Input lists must have the same length.\n\torders_vec = {}\n\trdp_vec = {}\n"
                         .format(orders_vec, rdp_vec))
    eps = rdp_vec - math.log(delta) / (orders_vec - 1)
    if np.isnan(eps).all():
        return np.inf, np.nan
    i = int(np.nanargmin(eps))
    return float(eps[i]), float(orders_vec[i])


class RDPIncrementalAccountant:
    """Cache of per-step RDP so per-round accounting is O(#orders) instead of recomputing the series
    for all 72 orders every round (the reference recomputes, ``privacy/__init__.py:232-233``)."""

    def __init__(self, orders):
        self.orders = list(orders)
        self._cache = {}

    def epsilon(self, q, sigma, steps, delta):
        key = (round(float(q), 12), round(float(sigma), 12))
        if key not in self._cache:
            self._cache[key] = np.array([_compute_rdp(q, sigma, o) for o in self.orders])
        return get_privacy_spent(self.orders, self._cache[key] * steps, delta)
