import numpy as np

from msrflute_b200.core.dataset import BaseDataset
from experiments.semisupervision.dataloaders.cifar_dataset import CIFAR100


class Dataset(BaseDataset):
    """``user_idx``: −1 labelled / −2 unlabelled / −3 RandAugment-ed unlabelled population (whole structure, no sample
    access), ≥ 0 one user's samples, ``test_only`` all users concatenated."""

    def __init__(self, data, test_only=False, user_idx=0, **kwargs):
        self.test_only, self.user_idx = test_only, user_idx
        args = kwargs.get("args", None)
        self.user_list, self.user_data, self.user_data_label, self.num_samples = self.load_data(data, test_only, args)
        self.features, self.labels = np.zeros((0,)), np.zeros((0,))
        if user_idx is not None and user_idx >= 0 or test_only:
            if test_only:
                self.user = "test_only"
                self.features = np.concatenate([np.asarray(self.user_data[u]) for u in self.user_list])
                self.labels = np.concatenate([np.asarray(self.user_data_label[u]) for u in self.user_list])
            else:
                self.user = self.user_list[user_idx]
                self.features = np.asarray(self.user_data[self.user])
                self.labels = np.asarray(self.user_data_label[self.user])

    def __getitem__(self, idx):
        return np.asarray(self.features[idx], dtype=np.float32), int(self.labels[idx])

    def __len__(self):
        return len(self.features)

    def load_data(self, data, test_only, sup_config):
        if data is None or isinstance(data, str):
            data = CIFAR100(self.user_idx, test_only, sup_config).data
        return data["users"], data["user_data"], data["user_data_label"], data["num_samples"]
