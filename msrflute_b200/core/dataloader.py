"""Abstract dataloader (ref. ``core/dataloader.py:7-12``): a torch DataLoader
whose ``create_loader()`` returns the iterable to loop over."""
from abc import ABC

from torch.utils.data import DataLoader as _TorchDataLoader


class BaseDataLoader(ABC, _TorchDataLoader):
    def create_loader(self):
        return self
