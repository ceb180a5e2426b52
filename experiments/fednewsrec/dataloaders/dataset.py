"""FedNewsRec dataset: one FL user = one MIND user; a training sample = (clicked-title matrix [50, 30],
candidate-title matrix [5, 30]) → index of the positive; an evaluation sample = one impression
(ref. ``experiments/fednewsrec/dataloaders/dataset.py``)."""
import numpy as np

from msrflute_b200.core.dataset import BaseDataset
from experiments.fednewsrec.dataloaders.preprocess_mind import MIND, get_test_input, get_train_input

_CACHE = {}


def _mind(root, mode):
    key = (root, mode)
    if key not in _CACHE:
        _CACHE[key] = MIND(root, mode, seed=0 if mode == "train" else 1)
    return _CACHE[key]


def build_structure(root, mode, test_only):
    m = _mind(root, "train" if not test_only else "val")
    by_user = {}
    if not test_only:
        cands, users, labels = get_train_input(m.sessions, m.news_index)
        for c, u, l in zip(cands, users, labels):
            uid = m.sessions[u][0]
            by_user.setdefault(uid, {"x": [], "y": []})
            by_user[uid]["x"].append((m.news_title[m.user["click"][u]], m.news_title[c]))
            by_user[uid]["y"].append(int(l))
    else:
        for imp in get_test_input(m.sessions, m.news_index):
            uid = m.sessions[imp["user"]][0]
            by_user.setdefault(uid, {"x": [], "y": []})
            by_user[uid]["x"].append((m.news_title[m.user["click"][imp["user"]]], m.news_title[imp["docs"]]))
            by_user[uid]["y"].append(imp["labels"])
    users = sorted(by_user)
    return {"users": users, "num_samples": [len(by_user[u]["y"]) for u in users],
            "user_data": {u: by_user[u]["x"] for u in users}, "user_data_label": {u: by_user[u]["y"] for u in users}}


class Dataset(BaseDataset):
    def __init__(self, data, test_only=False, user_idx=0, **kwargs):
        self.test_only, self.user_idx = test_only, user_idx
        args = kwargs.get("args", {}) or {}
        self.user_list, self.user_data, self.user_data_label, self.num_samples = self.load_data(data, test_only, args)
        users = self.user_list if (test_only or user_idx == -1) else [self.user_list[user_idx]]
        self.user = "test_only" if test_only else self.user_list[user_idx]
        self.features = [x for u in users for x in self.user_data[u]]
        self.labels = [y for u in users for y in self.user_data_label[u]]

    def __getitem__(self, idx):
        click, cand = self.features[idx]
        return (np.asarray(click), np.asarray(cand)), self.labels[idx]

    def __len__(self):
        return len(self.features)

    def load_data(self, data, test_only, args=None):
        if data is None or isinstance(data, str):
            data = build_structure(data, "train", test_only)
        return data["users"], data["user_data"], data["user_data_label"], data["num_samples"]
