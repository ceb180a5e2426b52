"""MNIST (FedML json, 1000 users) dataset; synthetic stand-in of the same shape when ``data`` is None
(the reference downloads with ``wget`` instead, ``dataloaders/preprocessing.py:1-67``)."""
import numpy as np

from msrflute_b200.data import synthetic
from msrflute_b200.data.federated import ArrayFederatedDataset


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: synthetic.make_vector_classification(1000, 60, 784, 10, seed=1))
    synthetic_test = staticmethod(lambda: synthetic.make_vector_classification(1000, 8, 784, 10, seed=2))

    def transform(self, x):
        return np.asarray(x, dtype=np.float32)
