"""Synthetic federated datasets with the shapes of the reference's benchmark tasks.

There is no network in the build/bench environment, so every task can instantiate a synthetic stand-in of
the right shape and scale (BASELINE.md task table).  Samples are class-conditional (prototype + noise) so
models actually learn on them — tests assert the loss decreases.  Generation is deterministic in ``seed``.

All generators return the reference's data structure::

    {'users': [...], 'num_samples': [...], 'user_data': {user: array}, 'user_data_label': {user: array}}
"""
from __future__ import annotations

import numpy as np


def _struct():
    return {"users": [], "num_samples": [], "user_data": {}, "user_data_label": {}}


def _add(st, name, x, y):
    st["users"].append(name)
    st["num_samples"].append(len(x))
    st["user_data"][name] = x
    st["user_data_label"][name] = y


def _sizes(rng, num_users, mean, fixed):
    if fixed:
        return np.full(num_users, mean, dtype=np.int64)
    return np.maximum(2, rng.poisson(mean, size=num_users)).astype(np.int64)


def make_vector_classification(num_users=1000, mean_samples=60, dim=784, num_classes=10, seed=0, fixed=False,
                               dtype=np.float32, prefix="u", proto_seed=1234):
    """LR-MNIST-like: ``dim``-vectors in [0,1], ``num_classes`` labels (1000 users × ~60 in the benchmark).
    Class prototypes depend on ``proto_seed`` only, so train/val/test splits (different ``seed``) share the task."""
    rng = np.random.default_rng(seed)
    protos = np.random.default_rng(proto_seed).random((num_classes, dim)).astype(np.float32)
    st = _struct()
    for u, n in enumerate(_sizes(rng, num_users, mean_samples, fixed)):
        y = rng.integers(0, num_classes, size=n)
        x = np.clip(0.6 * protos[y] + 0.4 * rng.random((n, dim), dtype=np.float32), 0, 1).astype(dtype)
        _add(st, "{}{:05d}".format(prefix, u), x, y.astype(np.int64))
    return st


def make_image_classification(num_users=500, mean_samples=100, shape=(32, 32, 3), num_classes=100, seed=0,
                              fixed=True, dtype=np.uint8, prefix="u", scale=255.0, proto_seed=1234):
    """FedCIFAR-100-like (500 users × 100 × 32×32×3 uint8, HWC) or FEMNIST-like (``shape=(28,28)``, float)."""
    rng = np.random.default_rng(seed)
    protos = np.random.default_rng(proto_seed).random((num_classes,) + tuple(shape)).astype(np.float32)
    st = _struct()
    for u, n in enumerate(_sizes(rng, num_users, mean_samples, fixed)):
        y = rng.integers(0, num_classes, size=n)
        x = 0.5 * protos[y] + 0.5 * rng.random((n,) + tuple(shape), dtype=np.float32)
        x = (x * scale).astype(dtype) if np.issubdtype(dtype, np.integer) else (x * scale).astype(dtype)
        _add(st, "{}{:05d}".format(prefix, u), x, y.astype(np.int64))
    return st


def make_char_sequences(num_users=715, mean_samples=50, seq_len=80, vocab=90, seed=0, fixed=False, prefix="u",
                        proto_seed=1234):
    """FedShakespeare-like next-char prediction: x = tokens[t], y = tokens[t+1]; 0 is the pad id."""
    rng = np.random.default_rng(seed)
    trans = np.random.default_rng(proto_seed).dirichlet(np.full(vocab - 1, 0.05), size=vocab - 1)  # sparse Markov chain
    cdf = np.cumsum(trans, axis=1)
    st = _struct()
    for u, n in enumerate(_sizes(rng, num_users, mean_samples, fixed)):
        seq = np.empty((n, seq_len + 1), dtype=np.int64)
        seq[:, 0] = rng.integers(1, vocab, size=n)
        r = rng.random((n, seq_len))
        for t in range(seq_len):
            seq[:, t + 1] = 1 + (r[:, t, None] > cdf[seq[:, t] - 1]).sum(axis=1).clip(max=vocab - 2)
        _add(st, "{}{:05d}".format(prefix, u), seq[:, :-1].copy(), seq[:, 1:].copy())
    return st


def make_token_lists(num_users=100, mean_samples=20, max_len=25, vocab=10000, seed=0, prefix="u", as_text=False):
    """Reddit-like variable-length utterances (nlg_gru / mlm_bert).  ``user_data[u]`` is a list of id lists, or of
    space-separated pseudo-word strings when ``as_text``."""
    rng = np.random.default_rng(seed)
    zipf = 1.0 / np.arange(1, vocab + 1) ** 1.1
    zipf /= zipf.sum()
    st = {"users": [], "num_samples": [], "user_data": {}}
    for u, n in enumerate(_sizes(rng, num_users, mean_samples, False)):
        utts = []
        for _ in range(int(n)):
            L = int(rng.integers(3, max_len + 1))
            ids = rng.choice(vocab, size=L, p=zipf)
            utts.append(" ".join("w{}".format(i) for i in ids) if as_text else ids.tolist())
        name = "{}{:05d}".format(prefix, u)
        st["users"].append(name)
        st["num_samples"].append(len(utts))
        st["user_data"][name] = utts
    return st


def make_ecg(num_users=50, mean_samples=40, length=187, num_classes=5, seed=0, prefix="u"):
    """ECG-heartbeat-like 1-D signals (187 samples, 5 classes)."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1, length, dtype=np.float32)
    st = _struct()
    for u, n in enumerate(_sizes(rng, num_users, mean_samples, False)):
        y = rng.integers(0, num_classes, size=n)
        x = np.sin(2 * np.pi * (y[:, None] + 1) * t[None]) * np.exp(-3 * t)[None] + 0.1 * rng.standard_normal((n, length))
        _add(st, "{}{:05d}".format(prefix, u), x.astype(np.float32), y.astype(np.int64))
    return st
