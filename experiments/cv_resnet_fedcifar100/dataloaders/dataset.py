"""Fed-CIFAR-100 (500 train users × 100 images, 100 test users; uint8 HWC).  Samples are served as float32 CHW
with the reference's transform — ``astype(float32).T`` (ref ``dataloaders/dataset.py:36``: note ``.T`` of an HWC
image is C×W×H) and no normalisation.  Synthetic stand-in when ``data`` is None."""
import numpy as np
import torch

from msrflute_b200.data import synthetic
from msrflute_b200.data.federated import ArrayFederatedDataset


class Dataset(ArrayFederatedDataset):
    synthetic_train = staticmethod(lambda: synthetic.make_image_classification(500, 100, (32, 32, 3), 100, seed=5))
    synthetic_test = staticmethod(lambda: synthetic.make_image_classification(100, 100, (32, 32, 3), 100, seed=6))

    def transform(self, x):
        return np.asarray(x).astype(np.float32).T

    def transform_batch(self, x: torch.Tensor) -> torch.Tensor:
        return x.float().permute(0, 3, 2, 1)
