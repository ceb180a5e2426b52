"""Aggregation strategies.  ``select_strategy`` maps the YAML ``strategy`` string the way the reference
does (``core/strategies/__init__.py:9-22``): dga → DGA, fedavg/fedprox → FedAvg, fedlabels → FedLabels."""
from .base import BaseStrategy  # noqa: F401
from .fedavg import FedAvg
from .dga import DGA
from .fedlabels import FedLabels


def select_strategy(strategy):
    name = str(strategy).lower()
    if name == "dga":
        return DGA
    if name in ("fedavg", "fedprox"):
        return FedAvg
    if name == "fedlabels":
        return FedLabels
    raise ValueError(f"cannot use strategy type {strategy}")
