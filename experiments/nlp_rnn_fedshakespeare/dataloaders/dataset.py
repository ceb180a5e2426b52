"""FedShakespeare (715 users, sequences of 80 char ids, vocab 90, pad id 0); synthetic stand-in when ``data`` is None."""
import numpy as np
import torch

from msrflute_b200.data import synthetic
from msrflute_b200.data.federated import ArrayFederatedDataset


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: synthetic.make_char_sequences(715, 50, 80, 90, seed=7))
    synthetic_test = staticmethod(lambda: synthetic.make_char_sequences(100, 30, 80, 90, seed=8))

    def transform(self, x):
        return np.asarray(x, dtype=np.int64)

    def transform_batch(self, x: torch.Tensor) -> torch.Tensor:
        return x.long()
