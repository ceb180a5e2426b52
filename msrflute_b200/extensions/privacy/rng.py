"""Randomness for differential privacy.

DP noise must be unpredictable: it may NOT come from the reproducibility seed (``server_config.b200.seed`` seeds the
global python / numpy / torch generators identically on every rank, so noise drawn from them would be publicly
reproducible and identical across ranks — i.e. cancellable).  Everything in ``extensions/privacy`` and the fused
global-DP server update draws from here instead:

* :func:`dp_seed` — a fresh 62-bit Philox key from the OS entropy pool (``secrets``), one per call;
* :func:`dp_generator` — a per-process, per-device ``torch.Generator`` seeded once from the OS entropy pool.

Deterministic noise is an explicit opt-in for tests: set ``FLUTE_DP_TEST_SEED=<int>`` (or call
:func:`set_test_seed`); the rank is mixed in so ranks never share a noise stream, and a counter makes successive
draws differ.
"""
import os
import secrets

import torch

_TEST_SEED = None
_COUNTER = 0
_GENERATORS = {}


def _rank() -> int:
    return int(os.environ.get("RANK", 0))


def set_test_seed(seed):
    """Opt into deterministic DP noise (tests only).  ``None`` restores OS entropy."""
    global _TEST_SEED, _COUNTER
    _TEST_SEED = None if seed is None else int(seed)
    _COUNTER = 0
    _GENERATORS.clear()


def _test_seed():
    if _TEST_SEED is not None:
        return _TEST_SEED
    env = os.environ.get("FLUTE_DP_TEST_SEED")
    return int(env) if env not in (None, "") else None


def dp_seed(stream: int = 0) -> int:
    """62-bit seed for a Philox-keyed noise kernel.  ``stream`` distinguishes uses in deterministic test mode."""
    global _COUNTER
    ts = _test_seed()
    if ts is None:
        return secrets.randbits(62)
    _COUNTER += 1
    x = (ts * 0x9E3779B97F4A7C15 + (_rank() + 1) * 0xBF58476D1CE4E5B9 + stream * 0x94D049BB133111EB + _COUNTER)
    x ^= x >> 31
    return x & ((1 << 62) - 1)


def dp_generator(device="cpu") -> torch.Generator:
    """Process-wide generator for DP noise on ``device`` — never seeded from the reproducibility seed."""
    dev = torch.device(device)
    key = (dev.type, dev.index)
    g = _GENERATORS.get(key)
    if g is None:
        g = torch.Generator(device=dev)
        g.manual_seed(dp_seed(stream=hash(key) & 0xFFFF) & 0x7FFFFFFFFFFFFFFF)
        _GENERATORS[key] = g
    return g


def randn_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty_like(t).normal_(generator=dp_generator(t.device))


def randn(shape, device="cpu") -> torch.Tensor:
    return torch.empty(shape, device=device).normal_(generator=dp_generator(device))


def rand(shape=(), device="cpu") -> torch.Tensor:
    return torch.empty(shape, device=device).uniform_(generator=dp_generator(device))
