"""MIT-BIH-style heartbeat dataset: 187-sample signals, 5 classes (ref. ``experiments/ecg_cnn/dataloaders/dataset.py``
reads an HDF5 blob); accepts .hdf5 (if h5py is installed) / .json / .npz / in-memory structures, and generates a
synthetic stand-in when ``data`` is None."""
import numpy as np
import torch

from msrflute_b200.data import synthetic
from msrflute_b200.data.federated import ArrayFederatedDataset


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: synthetic.make_ecg(50, 40, seed=21))
    synthetic_test = staticmethod(lambda: synthetic.make_ecg(10, 40, seed=22))

    def transform(self, x):
        return np.asarray(x, dtype=np.float32).reshape(1, -1)          # (channels=1, length)

    def transform_batch(self, x: torch.Tensor) -> torch.Tensor:
        return x.float().unsqueeze(1)
