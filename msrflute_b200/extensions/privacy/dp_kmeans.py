"""Differentially-private k-means (ref. ``extensions/privacy/dp_kmeans.py``).

The reference monkey-patches scikit-learn's private Lloyd kernel (``dp_kmeans.py:169``) — fragile across
sklearn versions.  This is a self-contained implementation with the same mechanism and a sklearn-like
surface (``fit / predict / fit_predict / cluster_centers_ / eps``): sphere-packing initialisation
(ref :22-49), per-iteration clipping of sample norms and weights, Gaussian noise on the weighted cluster
sums and counts with σ = √(2 ln(1.25/δ))·√(L² + w²)/ε (ref :51-73), privacy loss ε per iteration.  The Lloyd
iterations run in PyTorch, on the GPU when one is available.
"""
import numpy as np
import torch
from scipy.special import gammainc

from . import rng


def sample(ndim, r, num_samples=1):
    """Uniform samples from the ``ndim``-ball of radius ``r``."""
    x = np.random.normal(size=(num_samples, ndim))
    ssq = np.sum(x ** 2, axis=1)
    fr = r * gammainc(ndim / 2, ssq / 2) ** (1 / ndim) / np.sqrt(ssq)
    return x * fr.reshape(num_samples, 1)


def sphere_packing_initialization(n_clusters, n_dim, min_cluster_radius, max_space_size, max_failed_cases,
                                  verbose=None):
    """Data-independent init: centres at least ``2a`` apart inside the ball; ``a`` is halved after
    ``max_failed_cases`` rejected draws."""
    a, max_r = min_cluster_radius, max_space_size
    centers = np.empty((n_clusters, n_dim))
    cid = fails = 0
    while cid < n_clusters:
        v = sample(n_dim, max_r - a)[0]
        if cid > 0 and np.min(np.linalg.norm(centers[:cid] - v, axis=-1)) < 2 * a:
            fails += 1
            if fails >= max_failed_cases:
                fails, cid, a = 0, 0, a / 2
                if verbose:
                    print(f"Failing to pack, halving min_cluster_radius to {a}")
            continue
        centers[cid] = v
        cid += 1
    return centers, a


def add_gaussian_noise(centers_new, weight_in_clusters, eps, max_cluster_l2, max_sample_weight,
                       cluster_to_weight_ratio=-1, delta=1e-7, verbose=None):
    """In-place Gaussian mechanism on (weighted centre sums, cluster weights)."""
    scaler = 1.0
    if cluster_to_weight_ratio > 0:
        scaler = max_cluster_l2 / (max_sample_weight * cluster_to_weight_ratio)
    msw = max_sample_weight * scaler
    sigma = np.sqrt(2 * np.log(1.25 / delta)) * np.sqrt(max_cluster_l2 ** 2 + msw ** 2) / eps
    sums = centers_new * weight_in_clusters.reshape(-1, 1) + sigma * rng.randn(tuple(centers_new.shape)).double().numpy()
    weight_in_clusters[:] = np.maximum(
        1e-10, weight_in_clusters * scaler + sigma * rng.randn(tuple(weight_in_clusters.shape)).double().numpy()) / scaler
    centers_new[:] = sums / weight_in_clusters.reshape(-1, 1)
    return sigma


class _DPKMeans:
    def __init__(self, n_dim, eps, max_cluster_l2, max_sample_weight, max_iter, cluster_to_weight_ratio, n_clusters,
                 tol, verbose, delta, init_centers, device=None):
        self.n_dim, self.eps_per_iter, self.max_cluster_l2 = n_dim, eps, max_cluster_l2
        self.max_sample_weight, self.max_iter = max_sample_weight, max_iter
        self.cluster_to_weight_ratio, self.n_clusters, self.tol = cluster_to_weight_ratio, n_clusters, tol
        self.verbose, self.delta = verbose, delta
        self.cluster_centers_ = np.array(init_centers, dtype=np.float64)
        self.eps = [0]
        self.n_iter_ = 0
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")

    def _assign(self, X, C):
        d = torch.cdist(X, C)
        return d.argmin(dim=1), d.min(dim=1).values

    def fit(self, X, y=None, sample_weight=None):
        X = torch.as_tensor(np.asarray(X), dtype=torch.float64, device=self.device).clone()
        n = X.shape[0]
        w = torch.ones(n, dtype=torch.float64, device=self.device) if sample_weight is None else \
            torch.as_tensor(np.asarray(sample_weight), dtype=torch.float64, device=self.device)
        w = w.clamp(max=self.max_sample_weight)
        norms = X.norm(dim=1)
        X = X * (self.max_cluster_l2 / norms.clamp(min=1e-30)).clamp(max=1.0).unsqueeze(1)
        C = torch.as_tensor(self.cluster_centers_, dtype=torch.float64, device=self.device)
        scale_tol = float(X.var(dim=0).mean()) * self.tol
        for it in range(self.max_iter):
            labels, _ = self._assign(X, C)
            onehot = torch.zeros(n, self.n_clusters, dtype=torch.float64, device=self.device)
            onehot.scatter_(1, labels.unsqueeze(1), w.unsqueeze(1))
            weight_in = onehot.sum(0)
            sums = onehot.t() @ X
            centers_new = torch.where(weight_in.unsqueeze(1) > 0, sums / weight_in.clamp(min=1e-30).unsqueeze(1), C)
            cn, wi = centers_new.cpu().numpy(), weight_in.cpu().numpy()
            add_gaussian_noise(cn, wi, self.eps_per_iter, self.max_cluster_l2, self.max_sample_weight,
                               self.cluster_to_weight_ratio, self.delta, self.verbose)
            C_new = torch.as_tensor(cn, dtype=torch.float64, device=self.device)
            shift = float(((C_new - C) ** 2).sum())
            C = C_new
            self.eps[0] += self.eps_per_iter
            self.n_iter_ = it + 1
            if shift <= scale_tol:
                break
        self.cluster_centers_ = C.cpu().numpy()
        labels, dist = self._assign(X, C)
        self.labels_ = labels.cpu().numpy()
        self.inertia_ = float((w * dist ** 2).sum())
        return self

    def predict(self, X, sample_weight=None):
        X = torch.as_tensor(np.asarray(X), dtype=torch.float64, device=self.device)
        C = torch.as_tensor(self.cluster_centers_, dtype=torch.float64, device=self.device)
        return self._assign(X, C)[0].cpu().numpy()

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_


def DPKMeans(n_dim, eps, max_cluster_l2, max_sample_weight=1.0, max_iter=300, cluster_to_weight_ratio=-1,
             n_clusters=8, tol=1e-4, verbose=0, delta=1e-7, max_failed_cases=300, min_cluster_radius=None,
             **kwargs):
    """Build a DP k-means estimator; total privacy loss ≤ ``eps × iterations`` (``estimator.eps[0]``)."""
    if min_cluster_radius is None:
        min_cluster_radius = max_cluster_l2 / n_clusters
    init, _ = sphere_packing_initialization(n_clusters, n_dim, min_cluster_radius, max_cluster_l2, max_failed_cases,
                                            verbose)
    return _DPKMeans(n_dim, eps, max_cluster_l2, max_sample_weight, max_iter, cluster_to_weight_ratio, n_clusters,
                     tol, verbose, delta, init, device=kwargs.get("device"))


def resetKMeans():
    """No-op: nothing global is patched here (kept for API parity)."""
