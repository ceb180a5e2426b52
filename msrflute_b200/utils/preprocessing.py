"""Offline data preparation: raw per-author text → the FLUTE "user blob"
``{users, num_samples, user_data[, user_data_label]}`` as JSON, HDF5 (when ``h5py`` is importable) or ``.npz``.

The reference ships three hard-coded scripts (``utils/preprocessing/{create-json,create-hdf5,from_json_to_hdf5}.py``:
read ``train.tsv`` with columns author/…/content/…, group unique ``content`` strings per ``author``, dump).  Same
behaviour here as functions + a CLI with real arguments::

    python -m msrflute_b200.utils.preprocessing tsv2json  train.tsv  train.json
    python -m msrflute_b200.utils.preprocessing tsv2hdf5  train.tsv  train.hdf5
    python -m msrflute_b200.utils.preprocessing json2hdf5 train.json train.hdf5      # or *.npz

De-duplication uses a per-user set (the reference tests ``content not in list`` per row: quadratic per user).
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys
from collections import OrderedDict

import numpy as np

REDDIT_COLUMNS = ["author", "num1", "content", "str1", "str2", "num2", "subreddit"]


def group_tsv_by_user(path, columns=None, user_col="author", text_col="content", delimiter="\t"):
    """Stream a delimited file and collect each user's unique texts, in first-seen order."""
    columns = list(columns or REDDIT_COLUMNS)
    ui, ti = columns.index(user_col), columns.index(text_col)
    texts, seen = OrderedDict(), {}
    csv.field_size_limit(min(sys.maxsize, 2 ** 31 - 1))
    with open(path, newline="", encoding="utf8") as f:
        for row in csv.reader(f, delimiter=delimiter, quoting=csv.QUOTE_NONE):
            if len(row) <= max(ui, ti):
                continue
            user, text = row[ui], row[ti]
            if user not in texts:
                texts[user], seen[user] = [], set()
            if text not in seen[user]:
                seen[user].add(text)
                texts[user].append(text)
    users = sorted(texts)                                   # pandas ``groupby`` order in the reference
    return {"users": users, "num_samples": [len(texts[u]) for u in users],
            "user_data": {u: {"x": texts[u]} for u in users}}


def write_json(blob, path):
    with open(path, "w", encoding="utf8") as f:
        json.dump(blob, f)


def read_json(path):
    with open(path, encoding="utf8") as f:
        return json.load(f)


def _user_x(entry):
    return entry["x"] if isinstance(entry, dict) and "x" in entry else entry


def write_hdf5(blob, path):
    """FLUTE HDF5 layout: datasets ``users``/``num_samples``; groups ``user_data/<user>/x`` (+ ``user_data_label``)."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is optional
        raise RuntimeError("h5py is not installed; write a .npz instead (write_npz) — every loader reads it") from e
    with h5py.File(path, "w") as f:
        f.create_dataset("users", data=[str(u).encode("utf8") for u in blob["users"]])
        f.create_dataset("num_samples", data=np.asarray(blob["num_samples"], dtype=np.int64))
        g = f.create_group("user_data")
        for u in blob["users"]:
            x = _user_x(blob["user_data"][u])
            if len(x) and isinstance(x[0], str):
                x = [s.encode("utf8") for s in x]
            g.create_group(str(u)).create_dataset("x", data=x)
        if blob.get("user_data_label"):
            gl = f.create_group("user_data_label")
            for u in blob["users"]:
                gl.create_dataset(str(u), data=np.asarray(blob["user_data_label"][u]))


def write_npz(blob, path):
    """Dependency-free container read by ``msrflute_b200.data.federated.load_structure``."""
    np.savez_compressed(path, users=np.asarray(blob["users"], dtype=object),
                        num_samples=np.asarray(blob["num_samples"], dtype=np.int64),
                        user_data=np.asarray({u: _user_x(v) for u, v in blob["user_data"].items()}, dtype=object),
                        user_data_label=np.asarray(blob.get("user_data_label") or {}, dtype=object))


def write_blob(blob, path):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".json":
        write_json(blob, path)
    elif ext in (".hdf5", ".h5"):
        write_hdf5(blob, path)
    elif ext == ".npz":
        write_npz(blob, path)
    else:
        raise ValueError("unknown output format {!r} (use .json, .hdf5 or .npz)".format(ext))


def reduce_users(blob, n_users):
    """First ``n_users`` users of a blob (what the reference's ``testing/create_data.py:24-33`` does with 25)."""
    users = list(blob["users"])[:n_users]
    out = {"users": users, "num_samples": list(blob["num_samples"])[:n_users],
           "user_data": {u: blob["user_data"][u] for u in users}}
    if blob.get("user_data_label"):
        out["user_data_label"] = {u: blob["user_data_label"][u] for u in users}
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("command", choices=["tsv2json", "tsv2hdf5", "json2hdf5", "convert"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--user-col", default="author")
    ap.add_argument("--text-col", default="content")
    ap.add_argument("--columns", default=",".join(REDDIT_COLUMNS))
    ap.add_argument("--max-users", type=int, default=None)
    a = ap.parse_args(argv)
    if a.command.startswith("tsv"):
        blob = group_tsv_by_user(a.src, a.columns.split(","), a.user_col, a.text_col)
    else:
        blob = read_json(a.src)
    if a.max_users:
        blob = reduce_users(blob, a.max_users)
    write_blob(blob, a.dst)
    print("wrote {}: {} users, {} samples".format(a.dst, len(blob["users"]), int(np.sum(blob["num_samples"]))))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
