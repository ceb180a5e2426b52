"""Dataset for the personalization scenario: ``user_data[u] = {'x': array}`` (ref. ``experiments/cv/dataloaders/dataset.py``)."""
import numpy as np

from msrflute_b200.core.dataset import BaseDataset
from experiments.cv.data import prepare_dataset

_CACHE = {}


class Dataset(BaseDataset):
    def __init__(self, data, test_only=False, user_idx=0, **kwargs):
        self.test_only, self.user_idx = test_only, user_idx
        self.user_list, self.user_data, self.user_data_label, self.num_samples = self.load_data(data, test_only)
        users = self.user_list if test_only else [self.user_list[user_idx]]
        self.user = "test_only" if test_only else self.user_list[user_idx]
        xs = [np.asarray(self.user_data[u]["x"] if isinstance(self.user_data[u], dict) else self.user_data[u]) for u in users]
        self.features = np.concatenate(xs) if xs else np.zeros((0,))
        self.labels = np.concatenate([np.asarray(self.user_data_label[u]) for u in users]) if users else np.zeros((0,))

    def __getitem__(self, idx):
        # Samples are stored (H, W, C); the model transposes (N, H, W, C) -> (N, C, W, H) itself.  (The reference stores
        # torchvision CHW tensors and returns ``.T`` = (W, H, C) here, ``dataset.py:21`` — same layout downstream.)
        return self.features[idx].astype(np.float32), self.labels[idx]

    def __len__(self):
        return len(self.features)

    def load_data(self, data, test_only):
        if data is None:
            if "d" not in _CACHE:
                _CACHE["d"] = prepare_dataset(rnd=2020, img_size=32, image="cifar", total_num_clients=100,
                                              image_path="./", alpha=1.0, wantTransform=False)
            data = _CACHE["d"][1] if test_only else _CACHE["d"][0]
        return data["users"], data["user_data"], data["user_data_label"], data["num_samples"]
