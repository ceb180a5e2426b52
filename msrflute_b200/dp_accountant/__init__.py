"""PRV privacy accountant — numerical composition of privacy-loss random variables (Gopi, Lee, Wutschitz,
"Numerical Composition of Differential Privacy", arXiv:2106.02848).  Capability parity with the reference's
``utils/dp-accountant/prv_accountant`` submodule: same public names, written from the paper.

    from msrflute_b200.dp_accountant import PRVAccountant, PoissonSubsampledGaussianMechanism
    prv = PoissonSubsampledGaussianMechanism(sampling_probability=256/60000, noise_multiplier=1.1)
    acc = PRVAccountant(prvs=[prv], max_self_compositions=[5000], eps_error=0.1, delta_error=1e-8)
    eps_low, eps_est, eps_up = acc.compute_epsilon(delta=1e-5, num_self_compositions=[4000])
"""
from .prv import (PrivacyRandomVariable, PrivacyRandomVariableTruncated, GaussianMechanism, LaplaceMechanism,  # noqa: F401
                  PureDPMechanism, PoissonSubsampledGaussianMechanism)
from .domain import Domain  # noqa: F401
from .accountant import (PRVAccountant, Accountant, DPSGDAccountant, RDP, find_noise_multiplier,  # noqa: F401
                         DiscretePrivacyRandomVariable, compute_safe_domain_size)
