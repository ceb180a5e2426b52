"""Composition engine + user-facing accountants (ref. ``prv_accountant/{accountant,composers,discretisers,
discrete_privacy_random_variable,dpsgd,other_accountants}.py``).

Pipeline: truncate each PRV to [−L, L] (L from an RDP tail bound, remark 5.6 of the paper) → discretise on an aligned
grid of mesh ε_err/√(k/2·log(12/δ_err)) by CDF differencing, shifting the grid so the discrete mean equals the true
mean → self-compose k times with one FFT power → convolve different mechanisms pairwise → read ε(δ) off the composed
distribution: δ(ε) = E[(1 − e^{ε−Y})₊]."""
import warnings
from typing import Optional, Sequence, Tuple, Union

import numpy as np
from scipy import optimize
from scipy.fft import irfft, rfft
from scipy.signal import convolve

from .domain import Domain
from .prv import PoissonSubsampledGaussianMechanism, PrivacyRandomVariable, PrivacyRandomVariableTruncated


class DiscretePrivacyRandomVariable:
    def __init__(self, pmf: np.ndarray, domain: Domain):
        self.pmf, self.domain = np.asarray(pmf, dtype=np.float64), domain

    def __len__(self):
        return len(self.pmf)

    def compute_epsilon_estimate(self, delta: float) -> float:
        if not 0 < delta < 1:
            raise ValueError("delta must be in (0, 1)")
        t = np.asarray(self.domain.ts(), dtype=np.float64)
        p = self.pmf
        d1 = np.flip(np.flip(p).cumsum())                      # Σ_{j≥i} p_j
        d2 = np.flip(np.flip(p * np.exp(-t)).cumsum())         # Σ_{j≥i} p_j e^{−t_j}
        ndelta = np.exp(t) * d2 - d1                           # −δ(t_i)
        if np.any(-ndelta > delta) is False or -ndelta[-1] > delta:
            raise RuntimeError("Cannot compute epsilon: domain too small for this delta")
        i = int(np.searchsorted(ndelta, -delta, side="left"))
        if i <= 0:
            raise RuntimeError("Cannot compute epsilon: delta exceeds the mass of the distribution")
        return float(np.log((d1[i] - delta) / d2[i]))

    def compute_epsilon(self, delta: float, delta_error: float, epsilon_error: float) -> Tuple[float, float, float]:
        return (self.compute_epsilon_estimate(delta + delta_error) - epsilon_error,
                self.compute_epsilon_estimate(delta),
                self.compute_epsilon_estimate(delta - delta_error) + epsilon_error)

    def compute_delta_estimate(self, epsilon: float) -> float:
        t = np.asarray(self.domain.ts(), dtype=np.float64)
        return float(np.where(t > epsilon, (1.0 - np.exp(epsilon) * np.exp(-t)) * self.pmf, 0.0).sum())


def discretise(prv: PrivacyRandomVariableTruncated, domain: Domain) -> DiscretePrivacyRandomVariable:
    """Cell-centred CDF differencing + mean-preserving grid shift."""
    t = np.asarray(domain.ts(), dtype=np.float64)
    half = domain.dt() / 2.0
    pmf = np.asarray(prv.cdf(t + half), dtype=np.float64) - np.asarray(prv.cdf(t - half), dtype=np.float64)
    pmf = np.maximum(pmf, 0.0)
    mean_d = float(np.dot(t, pmf))
    shift = prv.mean() - mean_d
    if not np.abs(shift) < half:
        raise RuntimeError("Discrete mean differs significantly from continuous mean.")
    return DiscretePrivacyRandomVariable(pmf, domain.shift_right(shift))


class _Composer:
    """Per-mechanism FFT self-composition, then a pairwise convolution tree across mechanisms."""

    def __init__(self, dprvs: Sequence[DiscretePrivacyRandomVariable]):
        self.dprvs = list(dprvs)

    @staticmethod
    def _power(d: DiscretePrivacyRandomVariable, n: int) -> DiscretePrivacyRandomVariable:
        if n == 1:
            return d
        size = len(d)
        base = d.domain.t_min() - d.domain.shifts()            # unshifted grid origin (multiple of dt, symmetric)
        i0 = int(np.round(-base / d.domain.dt()))              # index of t = 0 on the unshifted grid
        pmf = np.roll(d.pmf, -i0)
        out = irfft(rfft(pmf) ** n, n=size)
        out = np.roll(out, i0)
        dom = Domain(base, d.domain.t_max() - d.domain.shifts(), size).shift_right(d.domain.shifts() * n)
        return DiscretePrivacyRandomVariable(np.maximum(out, 0.0), dom)

    @staticmethod
    def _conv(a, b):
        size = len(a)
        base = a.domain.t_min() - a.domain.shifts()
        i0 = int(np.round(-base / a.domain.dt()))
        full = convolve(a.pmf, b.pmf, mode="full")             # full[k] ↔ t = 2·base + k·dt ; want t = base + j·dt
        pmf = full[i0:i0 + size]
        dom = Domain(base, a.domain.t_max() - a.domain.shifts(), size).shift_right(a.domain.shifts() + b.domain.shifts())
        return DiscretePrivacyRandomVariable(np.maximum(pmf, 0.0), dom)

    def compute_composition(self, num_self_compositions: Sequence[int]) -> DiscretePrivacyRandomVariable:
        parts = [self._power(d, int(n)) for d, n in zip(self.dprvs, num_self_compositions) if n > 0]
        while len(parts) > 1:
            nxt = [self._conv(parts[i], parts[i + 1]) for i in range(0, len(parts) - 1, 2)]
            if len(parts) % 2:
                nxt.append(parts[-1])
            parts = nxt
        return parts[0]


class RDP:
    """Rényi-DP accountant over the same PRV objects (baseline + domain-size heuristic; ref. ``other_accountants.py``)."""

    def __init__(self, prvs: Sequence[PrivacyRandomVariable]):
        self.orders = np.concatenate((np.linspace(1.01, 2, num=51), np.linspace(2, 20, num=100)[1:],
                                      np.linspace(20, 100, num=81)[1:]))
        self.rdps = [np.array([prv.rdp(float(a)) for a in self.orders]) for prv in prvs]

    def compute_epsilon(self, delta: float, num_self_compositions: Sequence[int]) -> Tuple[float, float, float]:
        rdp = sum(r * n for r, n in zip(self.rdps, num_self_compositions))
        eps = rdp - np.log(delta) / (self.orders - 1)
        e = float(np.nanmin(eps))
        return 0.0, e, e


def compute_safe_domain_size(prvs, max_self_compositions, eps_error, delta_error) -> float:
    total = sum(max_self_compositions)
    _, _, L_max = RDP(prvs).compute_epsilon(delta_error / 4, max_self_compositions)
    for prv in prvs:
        _, _, L = RDP([prv]).compute_epsilon(delta_error / 8 / total, [1])
        L_max = max(L_max, L)
    return max(L_max, eps_error) + 3


class PRVAccountant:
    def __init__(self, prvs: Union[PrivacyRandomVariable, Sequence[PrivacyRandomVariable]], eps_error: float,
                 delta_error: float, max_self_compositions: Sequence[int] = None, eps_max: Optional[float] = None):
        if isinstance(prvs, PrivacyRandomVariable):
            prvs = [prvs]
            if max_self_compositions is not None and not isinstance(max_self_compositions, (list, tuple)):
                max_self_compositions = [max_self_compositions]
        if max_self_compositions is None:
            max_self_compositions = [1] * len(prvs)
        if len(max_self_compositions) != len(prvs):
            raise ValueError("max_self_compositions must have one entry per PRV")
        self.eps_error, self.delta_error = eps_error, delta_error
        self.prvs, self.max_self_compositions = list(prvs), list(max_self_compositions)
        if eps_max is not None:
            L = eps_max
            warnings.warn(f"Assuming that true epsilon < {eps_max}. If this is not a valid assumption set `eps_max=None`.")
        else:
            L = compute_safe_domain_size(self.prvs, self.max_self_compositions, eps_error, delta_error)
        total = sum(self.max_self_compositions)
        mesh = eps_error / np.sqrt(total / 2 * np.log(12 / delta_error))
        domain = Domain.create_aligned(-L, L, mesh)
        tprvs = [PrivacyRandomVariableTruncated(p, domain.t_min(), domain.t_max()) for p in self.prvs]
        self.composer = _Composer([discretise(t, domain) for t in tprvs])

    def compute_composition(self, num_self_compositions) -> DiscretePrivacyRandomVariable:
        if num_self_compositions is None:
            num_self_compositions = [1] * len(self.prvs)
        if isinstance(num_self_compositions, (int, np.integer)):
            num_self_compositions = [int(num_self_compositions)]
        if (np.array(self.max_self_compositions) < np.array(num_self_compositions)).any():
            raise ValueError("Requested number of compositions exceeds the maximum number of compositions")
        return self.composer.compute_composition(num_self_compositions)

    def compute_delta(self, epsilon: float, num_self_compositions) -> Tuple[float, float, float]:
        f = self.compute_composition(num_self_compositions)
        return (float(f.compute_delta_estimate(epsilon + self.eps_error) - self.delta_error),
                float(f.compute_delta_estimate(epsilon)),
                float(f.compute_delta_estimate(epsilon - self.eps_error) + self.delta_error))

    def compute_epsilon(self, delta: float, num_self_compositions) -> Tuple[float, float, float]:
        return self.compute_composition(num_self_compositions).compute_epsilon(delta, self.delta_error, self.eps_error)


class Accountant:
    """Deprecated single-mechanism front end kept for backwards compatibility (DP-SGD only)."""

    def __init__(self, noise_multiplier: float, sampling_probability: float, delta: float, max_compositions: int,
                 eps_error: float = None, mesh_size: float = None, verbose: bool = False):
        warnings.warn("`Accountant` will be deprecated. Use `PRVAccountant` with `PoissonSubsampledGaussianMechanism`.",
                      DeprecationWarning)
        assert mesh_size is None
        self.delta = delta
        prv = PoissonSubsampledGaussianMechanism(sampling_probability=sampling_probability, noise_multiplier=noise_multiplier)
        self.accountant = PRVAccountant([prv], eps_error=eps_error, delta_error=delta / 1000,
                                        max_self_compositions=[max_compositions])

    def compute_epsilon(self, num_compositions: int):
        return self.accountant.compute_epsilon(self.delta, [num_compositions])


class DPSGDAccountant:
    def __init__(self, noise_multiplier: float, sampling_probability: float, max_steps: int, eps_error: float = 0.1,
                 delta_error: float = 1e-9):
        prv = PoissonSubsampledGaussianMechanism(noise_multiplier=noise_multiplier, sampling_probability=sampling_probability)
        self.accountant = PRVAccountant([prv], max_self_compositions=[max_steps], eps_error=eps_error, delta_error=delta_error)

    def compute_epsilon(self, delta: float, num_steps: int):
        return self.accountant.compute_epsilon(delta=delta, num_self_compositions=[num_steps])


def find_noise_multiplier(sampling_probability: float, num_steps: int, target_epsilon: float, target_delta: float,
                          eps_error: float = 0.1, mu_max: float = 100.0) -> float:
    """Smallest noise multiplier whose ε upper bound at ``target_delta`` after ``num_steps`` is ≤ ``target_epsilon``."""
    def eps_upper(mu):
        acc = DPSGDAccountant(mu, sampling_probability, num_steps, eps_error=eps_error, delta_error=target_delta / 1000)
        return acc.compute_epsilon(target_delta, num_steps)[2]

    hi = 1.0
    while eps_upper(hi) > target_epsilon:
        hi *= 2
        if hi > mu_max:
            raise RuntimeError("Couldn't find a noise multiplier below mu_max")
    lo = hi / 2
    while eps_upper(lo) < target_epsilon and lo > 1e-3:
        lo /= 2
    res = optimize.root_scalar(lambda mu: eps_upper(mu) - target_epsilon, bracket=[lo, hi], method="brentq",
                               xtol=eps_error / 100 if eps_error else 1e-3)
    mu = float(res.root)
    while eps_upper(mu) > target_epsilon:          # land on the private side of the root
        mu += eps_error / 100
    return mu
