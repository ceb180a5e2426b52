"""CIFAR-10 split evenly over synthetic users (ref. ``experiments/classif_cnn/dataloaders/cifar_dataset.py``:
1000 train / 200 test users).  Uses torchvision's local copy when present (no download is attempted — there may be
no network) and otherwise a synthetic CIFAR-shaped stand-in."""
import os

import numpy as np

from msrflute_b200.data import synthetic


def _process(data, targets, n_users):
    total = len(data)
    per = total // n_users
    users = ["{:04d}".format(u) for u in range(n_users)]
    return {"users": users, "num_samples": [per] * n_users,
            "user_data": {u: data[i * per:(i + 1) * per] for i, u in enumerate(users)},
            "user_data_label": {u: np.asarray(targets[i * per:(i + 1) * per]) for i, u in enumerate(users)}}


class CIFAR10:
    def __init__(self, root="./data", train_users=1000, test_users=200):
        try:
            import torchvision
            if not os.path.isdir(os.path.join(root, "cifar-10-batches-py")):
                raise FileNotFoundError
            tr = torchvision.datasets.CIFAR10(root=root, train=True, download=False)
            te = torchvision.datasets.CIFAR10(root=root, train=False, download=False)
            norm = lambda d: ((d.astype(np.float32) / 255.0) - 0.5) / 0.5
            self.trainset = _process(norm(tr.data), tr.targets, train_users)
            self.testset = _process(norm(te.data), te.targets, test_users)
        except Exception:
            self.trainset = synthetic.make_image_classification(train_users, 50, (32, 32, 3), 10, seed=31,
                                                                dtype=np.float32, scale=1.0)
            self.testset = synthetic.make_image_classification(test_users, 50, (32, 32, 3), 10, seed=32,
                                                               dtype=np.float32, scale=1.0)
