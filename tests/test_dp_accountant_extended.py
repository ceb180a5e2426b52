"""Second tier of PRV-accountant tests — the ground the reference submodule's suite covers
(``/root/reference/utils/dp-accountant/tests/{test_composer,test_domain,test_dpsgd,test_backwards_compatibility}.py`` and
``test_privacy_random_variables``), written against this package's own composition engine:
discretisation invariants, composer identities (power-of-one, FFT power == repeated convolution, convolution tree ==
closed form), privacy-random-variable CDF / mean / RDP properties, domain edge cases, DP-SGD inverse robustness, the
deprecated ``Accountant`` front end."""
import warnings

import numpy as np
import pytest
from scipy import stats

from msrflute_b200.dp_accountant import (Accountant, Domain, DPSGDAccountant, GaussianMechanism, LaplaceMechanism,
                                         PoissonSubsampledGaussianMechanism, PRVAccountant, PrivacyRandomVariableTruncated,
                                         PureDPMechanism, RDP, compute_safe_domain_size, find_noise_multiplier)
from msrflute_b200.dp_accountant.accountant import DiscretePrivacyRandomVariable, _Composer, discretise


def gauss_delta(eps, mu):
    return stats.norm.cdf(-eps / mu + mu / 2) - np.exp(eps) * stats.norm.cdf(-eps / mu - mu / 2)


def _disc(prv, L=12.0, dt=2e-3):
    dom = Domain.create_aligned(-L, L, dt)
    return discretise(PrivacyRandomVariableTruncated(prv, dom.t_min(), dom.t_max()), dom)


# ------------------------------------------------------------------------------------------------ discretisation
@pytest.mark.parametrize("prv", [GaussianMechanism(2.0), GaussianMechanism(0.9),
                                 PoissonSubsampledGaussianMechanism(sampling_probability=0.02, noise_multiplier=1.1)])
def test_discretisation_is_a_probability_vector_with_the_true_mean(prv):
    d = _disc(prv)
    assert np.all(d.pmf >= 0) and abs(d.pmf.sum() - 1.0) < 1e-9
    t = np.asarray(d.domain.ts(), dtype=np.float64)
    trunc = PrivacyRandomVariableTruncated(prv, -12.0, 12.0)
    assert abs(float(np.dot(t, d.pmf)) - trunc.mean()) < 1e-6            # grid shifted so the means agree exactly
    assert abs(d.domain.shifts()) < d.domain.dt() / 2


def test_laplace_atoms_survive_discretisation_through_the_accountant():
    """The Laplace privacy loss has point masses at +-mu: eps(delta) of one release can never exceed mu."""
    acc = PRVAccountant([LaplaceMechanism(0.7)], max_self_compositions=[1], eps_error=0.01, delta_error=1e-10)
    lo, est, up = acc.compute_epsilon(1e-9, [1])
    assert lo <= est <= up and est <= 0.7 + 0.02
    assert acc.compute_delta(0.7 + 0.02, [1])[1] < 1e-8


# ------------------------------------------------------------------------------------------------------ composer
def test_composing_once_is_the_identity():
    d = _disc(GaussianMechanism(3.0))
    out = _Composer([d]).compute_composition([1])
    assert out.domain == d.domain and np.array_equal(out.pmf, d.pmf)


def test_fft_power_equals_repeated_convolution():
    d = _disc(GaussianMechanism(3.0), L=8.0, dt=4e-3)
    p4 = _Composer([d]).compute_composition([4])
    c2 = _Composer._conv(d, d)
    c4 = _Composer._conv(c2, c2)
    assert abs(p4.domain.shifts() - c4.domain.shifts()) < 1e-12
    assert np.abs(p4.pmf - c4.pmf).max() < 1e-10


@pytest.mark.parametrize("sigmas,counts", [([3.0], [9]), ([3.0, 5.0], [4, 10]), ([2.0, 4.0, 8.0], [1, 3, 5])])
def test_composition_tree_matches_gaussian_closed_form(sigmas, counts):
    ds = [_disc(GaussianMechanism(s)) for s in sigmas]
    f = _Composer(ds).compute_composition(counts)
    mu = np.sqrt(sum(n / s ** 2 for s, n in zip(sigmas, counts)))
    for eps in (0.25, 1.0, 2.0):
        assert abs(f.compute_delta_estimate(eps) - gauss_delta(eps, mu)) < 2e-4
    assert abs(f.pmf.sum() - 1.0) < 1e-6


def test_zero_count_mechanisms_are_skipped():
    a, b = _disc(GaussianMechanism(3.0)), _disc(GaussianMechanism(1.0))
    only_a = _Composer([a, b]).compute_composition([5, 0])
    ref = _Composer([a]).compute_composition([5])
    assert np.allclose(only_a.pmf, ref.pmf)


def test_epsilon_delta_round_trip_on_a_composed_variable():
    f = _Composer([_disc(GaussianMechanism(4.0))]).compute_composition([16])
    for delta in (1e-3, 1e-5, 1e-7):
        eps = f.compute_epsilon_estimate(delta)
        assert abs(f.compute_delta_estimate(eps) - delta) < 0.02 * delta
    with pytest.raises(ValueError):
        f.compute_epsilon_estimate(0.0)


def test_small_delta_outside_the_domain_is_an_error():
    acc = PRVAccountant([GaussianMechanism(20.0)], max_self_compositions=[1], eps_error=1.0, delta_error=1e-3)
    with pytest.raises((RuntimeError, ValueError)):
        acc.compute_epsilon(1e-300, [1])


# ------------------------------------------------------------------------------------- privacy random variables
@pytest.mark.parametrize("prv", [GaussianMechanism(1.5), LaplaceMechanism(1.0), PureDPMechanism(0.3),
                                 PoissonSubsampledGaussianMechanism(sampling_probability=0.1, noise_multiplier=0.8)])
def test_cdf_is_monotone_between_zero_and_one(prv):
    t = np.linspace(-30, 30, 4001)
    c = np.asarray(prv.cdf(t), dtype=np.float64)
    assert c[0] < 1e-9 and c[-1] > 1 - 1e-9
    assert np.all(np.diff(c) >= -1e-12)


def test_gaussian_prv_is_the_shifted_normal():
    sigma = 2.5
    prv = GaussianMechanism(sigma)
    mu = 1.0 / sigma
    t = np.linspace(-3, 3, 31)
    assert np.allclose(prv.cdf(t), stats.norm.cdf((t - mu * mu / 2) / mu), atol=1e-12)
    assert abs(prv.mean() - mu * mu / 2) < 1e-12
    assert abs(prv.rdp(8.0) - 8.0 / (2 * sigma ** 2)) < 1e-12


def test_pure_dp_prv_is_supported_on_plus_minus_eps():
    prv = PureDPMechanism(0.4)
    assert prv.cdf(-0.4 - 1e-9) == 0.0 and prv.cdf(0.4 + 1e-9) == 1.0
    mass_hi = 1.0 - float(prv.cdf(0.0))
    assert abs(mass_hi - np.exp(0.4) / (1 + np.exp(0.4))) < 1e-12
    assert abs(prv.mean() - 0.4 * np.tanh(0.2)) < 1e-12


def test_subsampling_amplifies_privacy_and_q_equal_one_is_the_gaussian():
    sigma = 1.2
    full = PRVAccountant([GaussianMechanism(sigma)], max_self_compositions=[10], eps_error=0.05, delta_error=1e-9)
    q1 = PRVAccountant([PoissonSubsampledGaussianMechanism(sampling_probability=1.0, noise_multiplier=sigma)],
                       max_self_compositions=[10], eps_error=0.05, delta_error=1e-9)
    sub = PRVAccountant([PoissonSubsampledGaussianMechanism(sampling_probability=0.05, noise_multiplier=sigma)],
                        max_self_compositions=[10], eps_error=0.05, delta_error=1e-9)
    e_full, e_q1, e_sub = (a.compute_epsilon(1e-5, [10])[1] for a in (full, q1, sub))
    assert abs(e_full - e_q1) < 0.02
    assert e_sub < 0.2 * e_full


def test_truncated_prv_renormalises_and_clips():
    base = GaussianMechanism(1.0)
    tr = PrivacyRandomVariableTruncated(base, -1.0, 2.0)
    assert tr.cdf(-1.0 - 1e-9) == 0.0 and tr.cdf(2.0 + 1e-9) == 1.0
    mid = float(tr.cdf(0.5))
    want = (base.cdf(0.5) - base.cdf(-1.0)) / (base.cdf(2.0) - base.cdf(-1.0))
    assert abs(mid - want) < 1e-12
    assert -1.0 < tr.mean() < 2.0


# ---------------------------------------------------------------------------------------------------------- domain
def test_domain_ts_equality_and_shifts():
    d = Domain.create_aligned(-0.5, 0.75, 0.125)
    ts = np.asarray(d.ts(), dtype=np.float64)
    assert len(ts) == d.size() and abs(ts[1] - ts[0] - d.dt()) < 1e-15
    assert ts[0] <= -0.5 and ts[-1] >= 0.75
    assert Domain(d.t_min(), d.t_max(), d.size()) == d
    assert Domain(d.t_min(), d.t_max() + d.dt() * 2, d.size() + 2) != d
    r = d.shift_right(0.03)
    assert abs(r.t_min() - d.t_min() - 0.03) < 1e-15 and abs(r.shift_left(0.03).t_min() - d.t_min()) < 1e-15
    assert abs(r.dt() - d.dt()) < 1e-15


@pytest.mark.parametrize("dt", [0.1, 0.3, 1e-3, 7e-4])
def test_create_aligned_floating_point_edge_cases(dt):
    d = Domain.create_aligned(-3 * dt, 3 * dt, dt)                         # bounds that are exact multiples of dt
    k = d.t_min() / d.dt()
    assert abs(k - round(k)) < 1e-6 and d.size() % 2 == 0
    assert d.t_min() <= -3 * dt + 1e-12 and d.t_max() >= 3 * dt - 1e-12
    assert abs(d.dt() - dt) < 1e-12


# ------------------------------------------------------------------------------------------- DP-SGD front ends
def test_dpsgd_epsilon_is_in_a_sensible_range_and_monotone_in_steps():
    acc = DPSGDAccountant(noise_multiplier=0.8, sampling_probability=5e-3, max_steps=10_000, eps_error=0.1)
    e1 = acc.compute_epsilon(1e-6, 1_000)
    e2 = acc.compute_epsilon(1e-6, 10_000)
    assert 0 < e1[0] <= e1[1] <= e1[2] < e2[1] < 20
    assert e1[2] - e1[0] < 0.5                                             # the bounds bracket tightly


@pytest.mark.parametrize("q,steps,eps,delta", [(4e-3, 2000, 4.0, 1e-7), (2e-2, 300, 1.5, 1e-5), (0.3, 20, 8.0, 1e-6)])
def test_find_noise_multiplier_is_robust_over_regimes(q, steps, eps, delta):
    mu = find_noise_multiplier(sampling_probability=q, num_steps=steps, target_epsilon=eps, target_delta=delta, eps_error=0.1)
    got = DPSGDAccountant(mu, q, steps, eps_error=0.1, delta_error=delta / 1000).compute_epsilon(delta, steps)
    assert got[2] <= eps + 1e-6 and got[2] > eps - 0.35
    looser = DPSGDAccountant(mu * 0.9, q, steps, eps_error=0.1, delta_error=delta / 1000).compute_epsilon(delta, steps)
    assert looser[2] > eps                                                 # less noise would violate the target


def test_find_noise_multiplier_gives_up_above_mu_max():
    with pytest.raises(RuntimeError):
        find_noise_multiplier(sampling_probability=0.5, num_steps=500, target_epsilon=1e-3, target_delta=1e-9, eps_error=0.1,
                              mu_max=4.0)


def test_legacy_accountant_warns_and_agrees_with_the_new_front_end():
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        old = Accountant(noise_multiplier=1.0, sampling_probability=0.01, delta=1e-6, max_compositions=800, eps_error=0.1)
    assert any(issubclass(x.category, DeprecationWarning) for x in w)
    new = DPSGDAccountant(1.0, 0.01, 800, eps_error=0.1, delta_error=1e-9)
    assert abs(old.compute_epsilon(400)[1] - new.compute_epsilon(1e-6, 400)[1]) < 1e-3


def test_safe_domain_grows_with_compositions_and_rdp_is_an_upper_bound():
    prv = PoissonSubsampledGaussianMechanism(sampling_probability=0.01, noise_multiplier=1.0)
    l1 = compute_safe_domain_size([prv], [100], eps_error=0.1, delta_error=1e-9)
    l2 = compute_safe_domain_size([prv], [5000], eps_error=0.1, delta_error=1e-9)
    assert 0 < l1 < l2
    acc = PRVAccountant([prv], max_self_compositions=[5000], eps_error=0.1, delta_error=1e-9)
    for n in (100, 1000, 5000):
        assert acc.compute_epsilon(1e-6, [n])[2] <= RDP([prv]).compute_epsilon(1e-6, [n])[1] + 1e-9
