"""Privacy-loss random variables with closed-form CDFs (ref. ``prv_accountant/privacy_random_variables/*``).

For a mechanism with output distributions P (with the record) and Q (without), the PRV is Y = log(P(t)/Q(t)), t ~ P.
Each class exposes ``cdf(t)``, ``mean()`` and ``rdp(alpha)`` (for the domain-size heuristic)."""
from abc import ABC, abstractmethod

import numpy as np
from scipy import integrate, stats
from scipy.special import erfc


def _log_ndtr_safe(x):
    return stats.norm.logcdf(x)


class PrivacyRandomVariable(ABC):
    @abstractmethod
    def cdf(self, t):
        ...

    def mean(self) -> float:
        lo, hi = -50.0, 50.0
        pts = np.concatenate([[lo], -np.logspace(-5, 1, 40)[::-1], [0], np.logspace(-5, 1, 40), [hi]])
        m = 0.0
        for a, b in zip(pts[:-1], pts[1:]):
            mid = 0.5 * (a + b)
            m += mid * (self.cdf(b) - self.cdf(a))
        return float(m)

    def rdp(self, alpha: float) -> float:
        raise NotImplementedError


class PrivacyRandomVariableTruncated:
    """Y conditioned on [t_min, t_max] (mass outside is removed and the rest renormalised)."""

    def __init__(self, prv, t_min: float, t_max: float):
        self.prv, self.t_min, self.t_max = prv, t_min, t_max
        self.remaining_mass = float(prv.cdf(t_max) - prv.cdf(t_min))

    def cdf(self, t):
        t = np.asarray(t, dtype=np.float64)
        c = (self.prv.cdf(np.clip(t, self.t_min, self.t_max)) - self.prv.cdf(self.t_min)) / self.remaining_mass
        return np.where(t < self.t_min, 0.0, np.where(t >= self.t_max, 1.0, c))

    def mean(self) -> float:
        # E[Y] = t_max − ∫ CDF over the truncated support, integrated on a grid refined around 0
        # point masses (Laplace / pure-DP losses) are bracketed by a 2e-12 wide cell so the midpoint rule places them exactly
        atoms = np.asarray(getattr(self.prv, "atoms", lambda: [])(), dtype=np.float64)
        around = np.concatenate([atoms - 1e-12, atoms + 1e-12]) if atoms.size else atoms
        pts = np.unique(np.clip(np.concatenate([[self.t_min], -np.logspace(-5, np.log10(max(-self.t_min, 1e-4)), 200)[::-1], [0.0],
                                                np.logspace(-5, np.log10(max(self.t_max, 1e-4)), 200), [self.t_max], around]),
                                self.t_min, self.t_max))
        m = 0.0
        for a, b in zip(pts[:-1], pts[1:]):
            m += 0.5 * (a + b) * float(self.cdf(b) - self.cdf(a))
        return m


class GaussianMechanism(PrivacyRandomVariable):
    """N(μ, 2μ) with μ = 1/(2σ²): the PRV of the Gaussian mechanism with sensitivity 1."""

    def __init__(self, noise_multiplier: float):
        self.noise_multiplier = float(noise_multiplier)
        self.mu = 1.0 / (2.0 * self.noise_multiplier ** 2)

    def cdf(self, t):
        return stats.norm.cdf(np.asarray(t, dtype=np.float64), loc=self.mu, scale=np.sqrt(2 * self.mu))

    def mean(self):
        return self.mu

    def rdp(self, alpha):
        return alpha / (2.0 * self.noise_multiplier ** 2)


class LaplaceMechanism(PrivacyRandomVariable):
    """PRV of the Laplace mechanism with scale b = 1/μ·… parameterised like the reference by ``mu`` = sensitivity/scale."""

    def __init__(self, mu: float):
        self.mu = float(mu)

    def cdf(self, t):
        t = np.asarray(t, dtype=np.float64)
        mid = 0.5 * np.exp(0.5 * (np.clip(t, -self.mu, self.mu) - self.mu))
        return np.where(t >= self.mu, 1.0, np.where(t <= -self.mu, 0.0, mid))

    def atoms(self):
        return [-self.mu, self.mu]

    def rdp(self, alpha):
        mu = self.mu
        if alpha == 1:
            return mu + np.exp(-mu) - 1
        return 1.0 / (alpha - 1) * np.log(alpha / (2 * alpha - 1) * np.exp((alpha - 1) * mu)
                                          + (alpha - 1) / (2 * alpha - 1) * np.exp(-alpha * mu))


class PureDPMechanism(PrivacyRandomVariable):
    """Worst-case PRV of an ε-DP mechanism: Y = ε w.p. e^ε/(1+e^ε), −ε otherwise."""

    def __init__(self, eps: float):
        self.eps = float(eps)

    def cdf(self, t):
        t = np.asarray(t, dtype=np.float64)
        return np.where(t < -self.eps, 0.0, np.where(t < self.eps, 1.0 / (1.0 + np.exp(self.eps)), 1.0))

    def atoms(self):
        return [-self.eps, self.eps]

    def mean(self):
        e = self.eps
        return e * (np.exp(e) - 1) / (np.exp(e) + 1)

    def rdp(self, alpha):
        e = self.eps
        return 1.0 / (alpha - 1) * np.log(np.exp(alpha * e) / (1 + np.exp(e)) * (1 + np.exp(-(2 * alpha - 1) * e))) \
            if alpha != 1 else self.mean()


class PoissonSubsampledGaussianMechanism(PrivacyRandomVariable):
    """DP-SGD step: Gaussian noise σ, Poisson sampling rate p (remove adjacency).
    P = (1−p)·N(0,σ²) + p·N(1,σ²), Q = N(0,σ²);  Y ≤ y  ⟺  t ≤ σ² log((e^y − (1−p))/p) + ½."""

    def __init__(self, sampling_probability: float, noise_multiplier: float):
        self.p = np.longdouble(sampling_probability)
        self.sigma = np.longdouble(noise_multiplier)
        self.sampling_probability, self.noise_multiplier = float(sampling_probability), float(noise_multiplier)

    def cdf(self, t):
        t = np.asarray(t, dtype=np.longdouble)
        p, s = self.p, self.sigma
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            z = np.log((p - 1) / p + np.exp(t) / p)                    # log((e^t − (1−p))/p)
            x = s * z + 0.5 / s                                        # t*/σ
            val = (1 - p) * 0.5 * erfc(-(x / np.sqrt(np.longdouble(2))).astype(np.float64)) \
                + p * 0.5 * erfc(-((x - 1 / s) / np.sqrt(np.longdouble(2))).astype(np.float64))
        return np.where(t > np.log(1 - p), val, 0.0).astype(np.float64)

    def rdp(self, alpha):
        from ..extensions.privacy.analysis import _compute_rdp
        return _compute_rdp(self.sampling_probability, self.noise_multiplier, alpha)
