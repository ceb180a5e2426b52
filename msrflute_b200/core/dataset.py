"""Abstract federated dataset (ref. ``core/dataset.py:7-27``, ``doc/sphinx/scenarios.rst:6-45``).

A concrete dataset exposes ``user_list`` (names), ``user_data`` (name → samples),
``num_samples`` (per user) and optionally ``user_data_label``; ``load_data``
accepts either a path or an already-built ``{'users','num_samples','user_data',
['user_data_label']}`` structure.
"""
from abc import ABC, abstractmethod

from torch.utils.data import Dataset as _TorchDataset


class BaseDataset(ABC, _TorchDataset):
    @abstractmethod
    def __init__(self, **kwargs):
        super().__init__()

    @abstractmethod
    def __getitem__(self, idx, **kwargs):
        """Return sample ``idx``."""

    @abstractmethod
    def __len__(self):
        """Number of samples."""

    @abstractmethod
    def load_data(self, **kwargs):
        """Read or instantiate the underlying data."""
