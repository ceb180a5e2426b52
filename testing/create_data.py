"""Materialise the data files of the four hello-world tasks (``--task nlg_gru|mlm_bert|classif_cnn|ecg_cnn``).

The reference's ``testing/create_data.py`` downloads LEAF-Reddit from Google Drive / CIFAR-10 from torchvision and
writes ``./data/<task>/{train,val,test}_data.*``.  This box has no network, so the same files (same user-blob
structure, same per-task shapes: sentences for nlg_gru, lists of sentences for mlm_bert, 50 users of 32x32x3 images
for classif_cnn, 187-sample heartbeats for ecg_cnn) are generated from ``msrflute_b200.data.synthetic``.  The
``hello_world_*.yaml`` configs run without these files too (``*_data: null`` picks the same generators in memory);
pass the written paths in the config to exercise the file loaders.

    python testing/create_data.py --task nlg_gru [--out testing/data] [--users 25]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msrflute_b200.data import synthetic  # noqa: E402
from msrflute_b200.utils.preprocessing import write_blob  # noqa: E402


def _plain(st):
    """numpy → lists so the blob is JSON-serialisable."""
    def conv(v):
        return v.tolist() if hasattr(v, "tolist") else v
    out = {"users": [str(u) for u in st["users"]], "num_samples": [int(n) for n in st["num_samples"]],
           "user_data": {str(u): conv(v) for u, v in st["user_data"].items()}}
    if st.get("user_data_label"):
        out["user_data_label"] = {str(u): conv(v) for u, v in st["user_data_label"].items()}
    return out


def text_task(task, out_dir, users, vocab_dir):
    for i, split in enumerate(("train_data", "val_data", "test_data")):
        st = synthetic.make_token_lists(num_users=users, mean_samples=12, max_len=25, vocab=500, seed=10 + i,
                                        as_text=True)
        blob = _plain(st)
        if task == "nlg_gru":            # one sentence string per sample, wrapped as {'x': [...]}
            blob["user_data"] = {u: {"x": [s if isinstance(s, str) else " ".join(map(str, s)) for s in v]}
                                 for u, v in blob["user_data"].items()}
            path = os.path.join(out_dir, split + ".json")
        else:                            # mlm_bert: plain list of sentences per user, ".txt" holding JSON (as upstream)
            blob["user_data"] = {u: [s if isinstance(s, str) else " ".join(map(str, s)) for s in v]
                                 for u, v in blob["user_data"].items()}
            path = os.path.join(out_dir, split + ".txt")
        blob.pop("user_data_label", None)
        with open(path, "w", encoding="utf8") as f:
            json.dump(blob, f)
        print("wrote", path)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "testing", "build_vocab.py"), "--data-dir", out_dir,
                           "--target-dir", vocab_dir])


def array_task(task, out_dir, users):
    for i, split in enumerate(("train_data", "test_data")):
        if task == "classif_cnn":
            st = synthetic.make_image_classification(users, 20, (32, 32, 3), 10, seed=20 + i)
        else:
            st = synthetic.make_ecg(num_users=users, mean_samples=21, length=187, num_classes=5, seed=30 + i)
        ext = ".hdf5"
        try:
            import h5py  # noqa: F401
        except ImportError:
            ext = ".npz"
        path = os.path.join(out_dir, split + ext)
        write_blob(_plain(st), path)
        print("wrote", path)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", required=True, choices=["nlg_gru", "mlm_bert", "classif_cnn", "ecg_cnn"])
    ap.add_argument("--out", default=os.path.join(ROOT, "testing", "data"))
    ap.add_argument("--users", type=int, default=25)
    a = ap.parse_args(argv)
    out_dir = os.path.join(a.out, a.task)
    os.makedirs(out_dir, exist_ok=True)
    if a.task in ("nlg_gru", "mlm_bert"):
        text_task(a.task, out_dir, a.users, os.path.join(a.out, "models"))
    else:
        array_task(a.task, out_dir, max(a.users, 50) if a.task == "classif_cnn" else a.users)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
