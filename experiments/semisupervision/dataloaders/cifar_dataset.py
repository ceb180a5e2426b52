"""Three aligned CIFAR-100 views for FedLabels (ref. ``experiments/semisupervision/dataloaders/cifar_dataset.py``):
``user_idx = -1`` labelled train split (or the test set), ``-2`` unlabelled split (weak augmentation), ``-3`` the SAME
unlabelled images under RandAugment.  Users come from a Dirichlet(α) label partition (``isclust: 0``), a clustered
partition (1) or a mixed one (2); each user's indices are split ``val_ratio`` (unlabelled) / ``train_ratio``
(labelled).  The reference materialises everything as JSON under ``./data``; this keeps the structures in memory and
uses a local torchvision CIFAR-100 copy if present, a synthetic CIFAR-shaped set otherwise (no downloads)."""
import os

import numpy as np
import torch

from experiments.semisupervision.dataloaders.RandAugment import RandAugment

MEAN = np.array([0.4914, 0.4822, 0.4465], dtype=np.float32).reshape(3, 1, 1)
STD = np.array([0.2023, 0.1994, 0.2010], dtype=np.float32).reshape(3, 1, 1)
_CACHE = {}


def _raw(train, n_synth, num_classes):
    try:
        import torchvision
        if os.path.isdir("./data/cifar-100-python"):
            ds = torchvision.datasets.CIFAR100("./data", train=train, download=False)
            return ds.data.transpose(0, 3, 1, 2).astype(np.float32) / 255.0, np.asarray(ds.targets)
    except Exception:
        pass
    rng = np.random.default_rng(60 + int(train))
    protos = np.random.default_rng(1234).random((num_classes, 3, 32, 32)).astype(np.float32)
    y = rng.integers(0, num_classes, size=n_synth)
    return (0.5 * protos[y] + 0.5 * rng.random((n_synth, 3, 32, 32), dtype=np.float32)), y


def dirichlet_partition(labels, n_users, alpha, num_classes, seed):
    rng = np.random.RandomState(seed)
    parts = [[] for _ in range(n_users)]
    for k in range(num_classes):
        idx = np.where(labels == k)[0]
        rng.shuffle(idx)
        cuts = (np.cumsum(rng.dirichlet(np.repeat(alpha, n_users))) * len(idx)).astype(int)[:-1]
        for p, chunk in zip(parts, np.split(idx, cuts)):
            p.extend(chunk.tolist())
    return parts


def clustered_partition(labels, n_users, mixed=False, seed=0):
    rng = np.random.RandomState(seed)
    order = np.argsort(labels, kind="stable")
    parts = [c.tolist() for c in np.array_split(order, n_users)]
    if mixed:
        pool = rng.permutation(len(labels))
        extra = np.array_split(pool[:len(labels) // 5], n_users)
        parts = [p + e.tolist() for p, e in zip(parts, extra)]
    return parts


def build(args, n_synth=4000):
    key = tuple(sorted((k, str(v)) for k, v in (args or {}).items() if k in ("isclust", "alpha", "ensize", "seed", "train_ratio", "val_ratio")))
    if key in _CACHE:
        return _CACHE[key]
    a = dict(isclust=0, alpha=0.1, ensize=100, seed=0, train_ratio=0.2, val_ratio=0.8, num_classes=100)
    a.update(args or {})
    X, Y = _raw(True, n_synth, a["num_classes"])
    Xt, Yt = _raw(False, max(n_synth // 5, 200), a["num_classes"])
    n_users = int(a["ensize"])
    if a["isclust"] == 1:
        parts = clustered_partition(Y, n_users)
    elif a["isclust"] == 2:
        parts = clustered_partition(Y, n_users, mixed=True, seed=a["seed"])
    else:
        parts = dirichlet_partition(Y, n_users, a["alpha"], a["num_classes"], a["seed"])
    norm = lambda x: (x - MEAN) / STD
    ra = RandAugment(1, 10)
    sets = {k: {"users": [], "num_samples": [], "user_data": {}, "user_data_label": {}} for k in ("lab", "unlab", "unlab_rand")}
    for u, ind in enumerate(parts):
        ind = np.asarray(ind, dtype=np.int64)
        n_val = int(a["val_ratio"] * len(ind))
        n_tr = max(int(a["train_ratio"] * len(ind)), 1)
        ul, lb = ind[:n_val], ind[n_val:n_val + n_tr]
        if len(lb) == 0 or len(ul) == 0:
            lb = ul = ind
        name = "{:04d}".format(u)
        rand = np.stack([ra(torch.from_numpy(X[i])).numpy() for i in ul]) if len(ul) else X[ul]
        for key_, xs, ys in (("lab", norm(X[lb]), Y[lb]), ("unlab", norm(X[ul]), Y[ul]), ("unlab_rand", norm(rand), Y[ul])):
            s = sets[key_]
            s["users"].append(name); s["num_samples"].append(len(ys)); s["user_data"][name] = xs; s["user_data_label"][name] = ys
    test = {"users": [], "num_samples": [], "user_data": {}, "user_data_label": {}}
    for i, chunk in enumerate(np.array_split(np.arange(len(Yt)), 10)):
        name = "{:04d}".format(i)
        test["users"].append(name); test["num_samples"].append(len(chunk))
        test["user_data"][name] = norm(Xt[chunk]); test["user_data_label"][name] = Yt[chunk]
    _CACHE[key] = (sets["lab"], sets["unlab"], sets["unlab_rand"], test)
    return _CACHE[key]


class CIFAR100:
    def __init__(self, user_idx=None, test_only=None, args=None, read_data=True):
        lab, unlab, unlab_rand, test = build(args)
        if test_only:
            self.data = test
        else:
            self.data = {-2: unlab, -3: unlab_rand}.get(user_idx, lab)
