"""``compute-dp-epsilon``: (ε lower, ε estimate, ε upper) of DP-SGD hyper-parameters from the PRV accountant and the
Rényi-DP accountant (same flags as the reference's ``utils/dp-accountant/bin/compute-dp-epsilon``; the optional
TF-privacy GDP row is replaced by the closed-form GDP/CLT estimate, no TensorFlow needed)."""
from __future__ import annotations

import argparse
import math
import sys

from scipy import optimize, stats

from . import PRVAccountant
from .accountant import RDP
from .prv import PoissonSubsampledGaussianMechanism


def arg_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Compute DP epsilon for a set of training hyper-params")
    p.add_argument("-p", "--sampling-probability", type=float, required=True,
                   help="probability of a user being sampled into a batch (often batch_size/num_samples)")
    p.add_argument("-s", "--noise-multiplier", type=float, required=True, help="noise std / clipping bound")
    p.add_argument("-i", "--num-compositions", type=int, required=True, help="number of DP-SGD steps")
    p.add_argument("-d", "--delta", type=float, required=True, help="target delta")
    p.add_argument("-v", "--verbose", action="store_true", default=None)
    p.add_argument("--fail-on-no-value", action="store_true", default=None,
                   help="raise instead of printing n/a when an accountant fails")
    return p


def gdp_epsilon(sampling_probability: float, noise_multiplier: float, steps: int, delta: float) -> float:
    """Gaussian-DP central-limit estimate (Bu et al. 2019): μ = p·sqrt(T·(e^{1/σ²} − 1)), then invert
    δ(ε) = Φ(−ε/μ + μ/2) − e^ε·Φ(−ε/μ − μ/2)."""
    mu = sampling_probability * math.sqrt(steps * (math.exp(1.0 / noise_multiplier ** 2) - 1.0))

    def delta_of(eps):
        return stats.norm.cdf(-eps / mu + mu / 2) - math.exp(eps) * stats.norm.cdf(-eps / mu - mu / 2)

    if delta_of(0.0) <= delta:
        return 0.0
    hi = 1.0
    while delta_of(hi) > delta and hi < 1e4:
        hi *= 2
    return float(optimize.brentq(lambda e: delta_of(e) - delta, 0.0, hi))


def main(argv=None) -> int:
    a = arg_parser().parse_args(argv)
    prv = PoissonSubsampledGaussianMechanism(sampling_probability=a.sampling_probability,
                                             noise_multiplier=a.noise_multiplier)
    methods = {}
    prv_acc = PRVAccountant(prvs=prv, max_self_compositions=a.num_compositions, eps_error=0.1,
                            delta_error=a.delta / 1000)
    methods["PRV Accountant"] = lambda n: prv_acc.compute_epsilon(delta=a.delta, num_self_compositions=n)
    rdp_acc = RDP(prvs=[prv])
    methods["RDP Accountant"] = lambda n: rdp_acc.compute_epsilon(delta=a.delta, num_self_compositions=[n])
    methods["GDP Accountant"] = lambda n: (0.0, gdp_epsilon(a.sampling_probability, a.noise_multiplier, n, a.delta),
                                           float("inf"))
    for name, fn in methods.items():
        try:
            lo, est, up = fn(a.num_compositions)
            print(f"{name}:\t\teps_lower = {lo:6.3} eps_estimate = {est:6.3}, eps_upper = {up:6.3} ")
        except Exception as e:  # noqa: BLE001 - mirror the reference: report n/a unless asked to fail
            if a.fail_on_no_value:
                raise
            if a.verbose:
                print(f"{name}: {type(e).__name__}: {e}", file=sys.stderr)
            print(f"{name}:\t\tn/a")
    return 0


if __name__ == "__main__":
    sys.exit(main())
