import numpy as np
import torch

from msrflute_b200.data.federated import ArrayFederatedDataset
from experiments.classif_cnn.dataloaders.cifar_dataset import CIFAR10

_CACHE = {}


def _cifar():
    if "d" not in _CACHE:
        _CACHE["d"] = CIFAR10()
    return _CACHE["d"]


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: _cifar().trainset)
    synthetic_test = staticmethod(lambda: _cifar().testset)

    def transform(self, x):
        return np.asarray(x).astype(np.float32).T            # HWC -> CWH like the reference (``.T``)

    def transform_batch(self, x: torch.Tensor) -> torch.Tensor:
        return x.float().permute(0, 3, 2, 1)
