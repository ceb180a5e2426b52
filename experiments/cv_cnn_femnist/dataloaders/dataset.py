"""Federated EMNIST (3400 users, 28×28, 62 classes); synthetic stand-in when ``data`` is None."""
import numpy as np

from msrflute_b200.data import synthetic
from msrflute_b200.data.federated import ArrayFederatedDataset


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: synthetic.make_image_classification(
        3400, 100, (28, 28), 62, seed=3, fixed=False, dtype=np.float32, scale=1.0))
    synthetic_test = staticmethod(lambda: synthetic.make_image_classification(
        340, 40, (28, 28), 62, seed=4, fixed=False, dtype=np.float32, scale=1.0))
