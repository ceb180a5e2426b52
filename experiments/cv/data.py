"""Non-IID partitioning of an image classification set over simulated clients (ref. ``experiments/cv/data.py``).

``DataPartitioner`` supports the reference's two modes:

* **Dirichlet** (``alpha``): for every class draw client proportions ~ Dir(α) (clients already holding ≥ N/n samples get
  0), repeat until every client has at least ``num_c`` samples (ref :118-156);
* **fixed label distribution** (``lab_distr``): reproduce given per-client class counts (scaled by ``ratio``), leftovers
  go to the last client (ref :75-116).

``prepare_dataset`` builds the FLUTE ``{'users','num_samples','user_data','user_data_label'}`` structures; it uses a
local torchvision CIFAR copy when there is one and a synthetic CIFAR-shaped set otherwise (no downloads).
"""
import os

import numpy as np
from numpy.random import RandomState


class _ArraySet:
    """Minimal (data, targets) container with the torchvision dataset surface the partitioner needs."""

    def __init__(self, data, targets):
        self.data, self.targets = data, list(targets)

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return self.data[i], self.targets[i]


class DataPartitioner:
    def __init__(self, data, sizes=None, rnd=0, alpha=0, num_c=10, dataset=None, lab_distr=None, ratio=1,
                 img_size=32, wantTrans=False):
        self.data, self.dataset = data, dataset
        self.total_num = len(sizes) if sizes is not None else len(lab_distr)
        self.img_size, self.wantTrans = img_size, wantTrans
        if lab_distr is not None:
            self.partitions, self.dat_stat = self._fixed(data, lab_distr, ratio, num_c)
        else:
            self.partitions, self.ratio, self.dat_stat, self.endat_size = self._dirichlet(data, sizes, alpha, num_c, rnd)

    def get_lab_distr(self):
        return self.dat_stat

    def return_partition(self, partition, flag="data", is_train_set=True):
        idx = self.partitions[partition]
        if flag != "data":
            return [self.data[i][1] for i in idx]
        mean = np.array([125.3, 123.0, 113.9], dtype=np.float32) / 255
        std = np.array([63.0, 62.1, 66.7], dtype=np.float32) / 255
        x = np.stack([np.asarray(self.data[i][0], dtype=np.float32) for i in idx]) if len(idx) else np.zeros((0, 32, 32, 3), np.float32)
        if x.size and x.max() > 1.5:
            x = x / 255.0
        x = (x - mean) / std
        if self.wantTrans:                       # per-client rotation (quarter turns keep this dependency-free)
            x = np.rot90(x, k=int(4 * partition / max(self.total_num, 1)) % 4, axes=(1, 2)).copy()
        return {"x": x}

    @staticmethod
    def _counts(labels, parts):
        out = {}
        for i, idx in enumerate(parts):
            u, c = np.unique(labels[idx], return_counts=True) if len(idx) else ([], [])
            out[i] = {int(a): int(b) for a, b in zip(u, c)}
        return out

    def _fixed(self, data, lab_distr, ratio, num_c):
        labels = np.array(data.targets)
        pool = {lab: np.where(labels == lab)[0] for lab in range(num_c)}
        parts = []
        keys = list(lab_distr.keys())
        for key in keys[:-1]:
            take = []
            for lab, num in lab_distr[key].items():
                k = min(int(num * ratio), len(pool[lab]))
                take += list(pool[lab][:k])
                pool[lab] = pool[lab][k:]
            parts.append(take)
        parts.append([i for lab in range(num_c) for i in pool[lab]])
        return parts, self._counts(labels, parts)

    def _dirichlet(self, data, psizes, alpha, num_c, rnd):
        n_nets, labels = len(psizes), np.array(data.targets)
        N = len(labels)
        rng = RandomState(rnd)
        min_size, guard = 0, 0
        while min_size < min(num_c, N // max(n_nets, 1)) and guard < 200:
            guard += 1
            batches = [[] for _ in range(n_nets)]
            for k in range(num_c):
                idx_k = np.where(labels == k)[0]
                rng.shuffle(idx_k)
                p = rng.dirichlet(np.repeat(alpha, n_nets))
                p = np.array([q * (len(b) < N / n_nets) for q, b in zip(p, batches)])
                p = p / p.sum() if p.sum() > 0 else np.full(n_nets, 1.0 / n_nets)
                cuts = (np.cumsum(p) * len(idx_k)).astype(int)[:-1]
                batches = [b + part.tolist() for b, part in zip(batches, np.split(idx_k, cuts))]
            min_size = min(len(b) for b in batches)
        for b in batches:
            rng.shuffle(b)
        sizes = np.array([len(b) for b in batches])
        return batches, sizes / sizes.sum(), self._counts(labels, batches), int(sizes.sum())


def _load_images(image, image_path, train):
    try:
        import torchvision
        root = os.path.join(image_path, image)
        cls = torchvision.datasets.CIFAR10 if image == "cifar" else torchvision.datasets.CIFAR100
        folder = "cifar-10-batches-py" if image == "cifar" else "cifar-100-python"
        if os.path.isdir(os.path.join(root, folder)):
            ds = cls(root=root, train=train, download=False)
            return _ArraySet(ds.data, ds.targets)
    except Exception:
        pass
    rng = np.random.default_rng(40 + int(train))
    nc = 10 if image == "cifar" else 100
    n = 5000 if train else 1000
    protos = rng.random((nc, 32, 32, 3), dtype=np.float32)
    y = rng.integers(0, nc, size=n)
    x = ((0.5 * protos[y] + 0.5 * rng.random((n, 32, 32, 3), dtype=np.float32)) * 255).astype(np.uint8)
    return _ArraySet(x, y.tolist())


def partition_dataset(rnd, img_size, image, total_num_clients, image_path, alpha, wantTransform):
    sizes = [1.0 / total_num_clients] * total_num_clients
    num_c = 10 if image == "cifar" else 100
    train = DataPartitioner(_load_images(image, image_path, True), sizes, rnd, alpha=alpha, num_c=num_c,
                            img_size=img_size, wantTrans=wantTransform)
    test = DataPartitioner(_load_images(image, image_path, False), sizes, rnd, alpha=alpha, num_c=num_c,
                           img_size=img_size, wantTrans=wantTransform)
    return train, test


def prepare_dataset(rnd=2020, img_size=32, image="cifar", total_num_clients=100, image_path="./", save_to_disk=False,
                    alpha=1.0, wantTransform=False):
    train, test = partition_dataset(rnd, img_size, image, total_num_clients, image_path, alpha, wantTransform)
    out = []
    for part, is_train in ((train, True), (test, False)):
        st = {"users": [], "num_samples": [], "user_data": {}, "user_data_label": {}}
        for c in range(total_num_clients):
            u = "{:04d}".format(c)
            x = part.return_partition(c, "data", is_train)
            y = part.return_partition(c, "label", is_train)
            st["users"].append(u)
            st["num_samples"].append(len(y))
            st["user_data"][u] = x
            st["user_data_label"][u] = np.asarray(y, dtype=np.int64)
        out.append(st)
    if save_to_disk:
        import json
        for name, st in zip(("train", "test"), out):
            with open(os.path.join(image_path, "{}_{}.json".format(image, name)), "w") as f:
                json.dump({k: (v if k in ("users", "num_samples") else {u: (np.asarray(a["x"] if isinstance(a, dict) else a).tolist())
                                                                         for u, a in v.items()}) for k, v in st.items()}, f)
    return out[0], out[1]
