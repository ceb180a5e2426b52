"""PRV accountant tests, modelled on the reference submodule's test ideas (``utils/dp-accountant/tests``): analytic
Gaussian δ bracketed by [δ_lower, δ_upper], homogeneous vs heterogeneous composition, invariance to
``max_self_compositions``, the ``find_noise_multiplier`` inverse property, error cases."""
import numpy as np
import pytest
from scipy import stats

from msrflute_b200.dp_accountant import (Domain, DPSGDAccountant, GaussianMechanism, LaplaceMechanism,
                                         PoissonSubsampledGaussianMechanism, PRVAccountant, PureDPMechanism, RDP,
                                         find_noise_multiplier)


def gauss_delta(eps, mu):
    return stats.norm.cdf(-eps / mu + mu / 2) - np.exp(eps) * stats.norm.cdf(-eps / mu - mu / 2)


@pytest.mark.parametrize("sigma,n,eps", [(8.0, 50, 1.0), (2.0, 4, 0.5), (20.0, 1000, 2.0)])
def test_gaussian_composition_matches_closed_form(sigma, n, eps):
    acc = PRVAccountant([GaussianMechanism(sigma)], max_self_compositions=[n], eps_error=0.05, delta_error=1e-9)
    lo, est, up = acc.compute_delta(eps, [n])
    true = gauss_delta(eps, np.sqrt(n) / sigma)
    assert lo <= true <= up
    assert abs(est - true) < 0.02 * true + 1e-9


def test_heterogeneous_equals_homogeneous_and_closed_form():
    a = PRVAccountant([GaussianMechanism(4.0), GaussianMechanism(8.0)], max_self_compositions=[10, 20], eps_error=0.05,
                      delta_error=1e-9)
    lo, est, up = a.compute_delta(1.0, [10, 20])
    true = gauss_delta(1.0, np.sqrt(10 / 16 + 20 / 64))
    assert lo <= true <= up
    b = PRVAccountant([GaussianMechanism(4.0), GaussianMechanism(4.0)], max_self_compositions=[5, 5], eps_error=0.05,
                      delta_error=1e-9)
    c = PRVAccountant([GaussianMechanism(4.0)], max_self_compositions=[10], eps_error=0.05, delta_error=1e-9)
    assert abs(b.compute_delta(1.0, [5, 5])[1] - c.compute_delta(1.0, [10])[1]) < 1e-4


def test_dpsgd_bounds_ordering_rdp_dominance_and_max_compositions_invariance():
    prv = PoissonSubsampledGaussianMechanism(sampling_probability=0.01, noise_multiplier=1.0)
    e1 = PRVAccountant([prv], max_self_compositions=[1000], eps_error=0.1, delta_error=1e-9).compute_epsilon(1e-6, [500])
    e2 = PRVAccountant([prv], max_self_compositions=[4000], eps_error=0.1, delta_error=1e-9).compute_epsilon(1e-6, [500])
    assert e1[0] <= e1[1] <= e1[2]
    assert abs(e1[1] - e2[1]) < 0.05                                   # estimate barely depends on the grid budget
    rdp_eps = RDP([prv]).compute_epsilon(1e-6, [500])[1]
    assert e1[2] < rdp_eps                                             # PRV accounting is tighter than RDP
    more = PRVAccountant([prv], max_self_compositions=[4000], eps_error=0.1, delta_error=1e-9).compute_epsilon(1e-6, [2000])
    assert more[1] > e2[1]


def test_requesting_more_than_max_compositions_is_an_error():
    acc = PRVAccountant([GaussianMechanism(4.0)], max_self_compositions=[10], eps_error=0.1, delta_error=1e-8)
    with pytest.raises(ValueError):
        acc.compute_epsilon(1e-5, [11])
    with pytest.raises(ValueError):
        PRVAccountant([GaussianMechanism(4.0)], max_self_compositions=[1, 2], eps_error=0.1, delta_error=1e-8)


def test_pure_dp_and_laplace_mechanisms():
    # k-fold composition of ε0-DP is at most k·ε0-DP: δ(k·ε0) must be ~0
    acc = PRVAccountant([PureDPMechanism(0.1)], max_self_compositions=[20], eps_error=0.05, delta_error=1e-9)
    assert acc.compute_delta(2.0 + 0.06, [20])[1] < 1e-6
    assert acc.compute_delta(0.5, [20])[1] > 0
    lap = PRVAccountant([LaplaceMechanism(0.5)], max_self_compositions=[8], eps_error=0.05, delta_error=1e-9)
    lo, est, up = lap.compute_epsilon(1e-4, [8])
    assert lo <= est <= up <= 8 * 0.5 + 0.1                            # never worse than basic composition


def test_find_noise_multiplier_inverts_the_accountant():
    mu = find_noise_multiplier(sampling_probability=0.05, num_steps=200, target_epsilon=3.0, target_delta=1e-5, eps_error=0.1)
    eps = DPSGDAccountant(mu, 0.05, 200, eps_error=0.1, delta_error=1e-8).compute_epsilon(1e-5, 200)
    assert eps[2] <= 3.0 + 1e-6 and eps[2] > 2.7


def test_domain_alignment():
    d = Domain.create_aligned(-1.03, 2.02, 0.01)
    assert d.size() % 2 == 0 and abs(d.dt() - 0.01) < 1e-10
    assert abs(d.t_min() / d.dt() - round(d.t_min() / d.dt())) < 1e-6      # 0 is a grid point
    s = d.shift_right(0.004)
    assert abs(s.shifts() - 0.004) < 1e-12 and s.size() == d.size()
    with pytest.raises(ValueError):
        Domain(0, 1, 11)
