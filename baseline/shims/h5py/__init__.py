"""Offline stand-in for ``h5py`` (not installed): ``File(path)`` opens an ``.npz`` written by
``baseline/run_reference.py`` and exposes the ``f['examples'][user]['image'][()]`` access pattern the
reference's Fed-CIFAR-100 / FEMNIST readers use."""
import numpy as np


class _Leaf:
    def __init__(self, arr):
        self._a = arr

    def __getitem__(self, k):
        return self._a if k == () else self._a[k]

    def __len__(self):
        return len(self._a)

    @property
    def shape(self):
        return self._a.shape


class _Group(dict):
    pass


class File(_Group):
    def __init__(self, path, mode="r", **kw):
        super().__init__()
        z = np.load(path, allow_pickle=True)
        ex = _Group()
        users = [str(u) for u in z["users"]]
        off = z["offsets"]
        arrays = {k[2:]: z[k] for k in z.files if k.startswith("f_")}      # each npz member is read ONCE
        for i, u in enumerate(users):
            g = _Group()
            for name, arr in arrays.items():
                g[name] = _Leaf(arr[off[i]:off[i + 1]])
            ex[u] = g
        self["examples"] = ex

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
