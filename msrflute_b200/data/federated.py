"""Reusable dataset / dataloader bases for array-shaped federated data.

Task plug-ins under ``experiments/<task>/dataloaders`` subclass these; the attribute contract is the
reference's (``core/dataset.py`` + ``doc/sphinx/scenarios.rst:6-45``): ``user_list``, ``user_data``,
``user_data_label``, ``num_samples``; ``test_only`` datasets concatenate all users; ``user_idx`` selects one.

B200 hook: ``device_tensors(user)`` returns the user's *fully transformed* samples as tensors so the
device-resident engine (``core/engine.py``) can keep every shard in HBM and build mini-batches with an
on-device gather instead of a python DataLoader.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from ..core.dataloader import BaseDataLoader
from ..core.dataset import BaseDataset


def load_structure(data):
    """Accept an in-memory structure, a ``.json`` / ``.pt`` / ``.npz`` path, or (with h5py) a FLUTE ``.hdf5``."""
    if isinstance(data, dict) or hasattr(data, "keys"):
        return data
    if isinstance(data, str):
        ext = os.path.splitext(data)[1].lower()
        if ext == ".json":
            with open(data) as f:
                return json.load(f)
        if ext in (".pt", ".pth"):
            return torch.load(data, weights_only=False)
        if ext == ".npz":
            z = np.load(data, allow_pickle=True)
            return {k: (z[k].item() if z[k].dtype == object and z[k].shape == () else z[k]) for k in z.files}
        if ext in (".hdf5", ".h5"):
            import h5py  # optional dependency
            out = {"users": [], "num_samples": [], "user_data": {}, "user_data_label": {}}
            with h5py.File(data, "r") as f:
                out["users"] = [u.decode() if isinstance(u, bytes) else str(u) for u in f["users"][()]]
                out["num_samples"] = [int(n) for n in f["num_samples"][()]]
                for u in out["users"]:
                    out["user_data"][u] = f["user_data"][u]["x"][()] if "x" in f["user_data"][u] else f["user_data"][u][()]
                    if "user_data_label" in f:
                        out["user_data_label"][u] = f["user_data_label"][u][()]
            return out
    raise ValueError("unsupported data source: {!r}".format(data))


class ArrayFederatedDataset(BaseDataset):
    """Features + labels per user."""

    #: subclasses: callable producing the synthetic structure when ``data`` is None
    synthetic_train = None
    synthetic_test = None

    def __init__(self, data, test_only=False, user_idx=0, **kwargs):
        self.test_only = test_only
        self.user_idx = user_idx
        self.args = kwargs.get("args", None)
        self.user_list, self.user_data, self.user_data_label, self.num_samples = self.load_data(data, test_only)
        if user_idx == -1 or test_only:
            self.user = "test_only"
            self.features = np.concatenate([np.asarray(self.user_data[u]) for u in self.user_list]) \
                if len(self.user_list) else np.zeros((0,))
            self.labels = np.concatenate([np.asarray(self.user_data_label[u]) for u in self.user_list]) \
                if len(self.user_list) else np.zeros((0,))
        else:
            if user_idx is None:
                raise ValueError("in train mode, user_idx must be specified")
            self.user = self.user_list[user_idx]
            self.features = self.user_data[self.user]
            self.labels = self.user_data_label[self.user]

    # -- overridable -----------------------------------------------------------
    def transform(self, x):
        return np.asarray(x, dtype=np.float32)

    def transform_batch(self, x: torch.Tensor) -> torch.Tensor:
        """Vectorised version of :meth:`transform` on a tensor of raw samples (device path)."""
        return x.float()

    def label_dtype(self):
        return np.int64

    def load_data(self, data, test_only):
        if data is None:
            gen = type(self).synthetic_test if test_only else type(self).synthetic_train
            if gen is None:
                raise ValueError("no data given and the task defines no synthetic generator")
            data = gen()
        data = load_structure(data)
        return data["users"], data["user_data"], data["user_data_label"], data["num_samples"]

    # -- torch dataset -----------------------------------------------------------
    def __getitem__(self, idx):
        return self.transform(self.features[idx]), self.labels[idx]

    def __len__(self):
        return len(self.features)

    # -- device hook ---------------------------------------------------------------
    def device_tensors(self, user):
        """Raw (untransformed) per-user tensors; ``transform_batch`` is applied after the on-device gather."""
        return {"x": torch.as_tensor(np.asarray(self.user_data[user])),
                "y": torch.as_tensor(np.asarray(self.user_data_label[user]).astype(self.label_dtype()))}


class ArrayDataLoader(BaseDataLoader):
    """``{'x', 'y'}`` batches; shuffles in train mode (ref. e.g. ``cv_resnet_fedcifar100/dataloaders/dataloader.py``)."""

    dataset_class = ArrayFederatedDataset

    def __init__(self, mode, num_workers=0, **kwargs):
        args = kwargs["args"]
        self.batch_size = args["batch_size"]
        self.mode = mode
        dataset = self.dataset_class(data=kwargs["data"], test_only=(mode != "train"),
                                     user_idx=kwargs.get("user_idx", None), args=args)
        super().__init__(dataset, batch_size=self.batch_size, shuffle=(mode == "train"), num_workers=num_workers,
                         collate_fn=self.collate_fn)

    @staticmethod
    def collate_fn(batch):
        x, y = zip(*batch)
        return {"x": torch.as_tensor(np.stack(x)), "y": torch.as_tensor(np.asarray(y))}
