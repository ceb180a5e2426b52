"""Offline tooling: data preparation, vocabulary builder, DP-epsilon CLI, timing helpers, async checkpoint writer."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_preprocessing_tsv_to_json_and_npz_roundtrip(tmp_path):
    from msrflute_b200.data.federated import load_structure
    from msrflute_b200.utils import preprocessing as P
    tsv = tmp_path / "t.tsv"
    tsv.write_text("bob\t1\thello world\ta\tb\t2\tsub\nalice\t1\tfoo bar\ta\tb\t2\tsub\n"
                   "bob\t1\thello world\ta\tb\t2\tsub\nbob\t1\tsecond post\ta\tb\t2\tsub\n")
    out = tmp_path / "o.json"
    assert P.main(["tsv2json", str(tsv), str(out)]) == 0
    blob = json.loads(out.read_text())
    assert blob["users"] == ["alice", "bob"] and blob["num_samples"] == [1, 2]        # duplicates dropped per user
    assert blob["user_data"]["bob"]["x"] == ["hello world", "second post"]
    npz = tmp_path / "o.npz"
    assert P.main(["convert", str(out), str(npz), "--max-users", "1"]) == 0
    z = load_structure(str(npz))
    assert list(z["users"]) == ["alice"] and list(z["user_data"]["alice"]) == ["foo bar"]


def test_create_data_and_build_vocab_scripts(tmp_path):
    out = str(tmp_path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "testing", "create_data.py"), "--task", "nlg_gru", "--out", out,
                        "--users", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    blob = json.load(open(os.path.join(out, "nlg_gru", "train_data.json")))
    assert len(blob["users"]) == 5 and all(isinstance(s, str) for s in blob["user_data"][blob["users"][0]]["x"])
    vocab = json.load(open(os.path.join(out, "models", "vocab_reddit.vocab")))
    assert vocab["vocab"]["<PAD>"] == vocab["pad_symbol"] == 0 and vocab["size"] == len(vocab["vocab"]) > 10


def test_compute_dp_epsilon_cli():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "compute-dp-epsilon"), "-p", "0.01", "-s", "1.0", "-i", "200",
                        "-d", "1e-5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = {l.split(":")[0]: l for l in r.stdout.splitlines() if "eps_estimate" in l}
    assert {"PRV Accountant", "RDP Accountant", "GDP Accountant"} <= set(rows)
    est = {k: float(v.split("eps_estimate =")[1].split(",")[0]) for k, v in rows.items()}
    assert est["PRV Accountant"] <= est["RDP Accountant"] + 1e-6                     # PRV is the tighter accountant
    assert 0.0 < est["PRV Accountant"] < 10.0


def test_round_timer_and_trace_window_are_inert_without_cuda():
    from msrflute_b200.utils.timing import RoundTimer, TraceWindow, nvtx_range
    t = RoundTimer()
    t.mark(0)
    assert t.drain() == [] or torch.cuda.is_available()
    with nvtx_range("phase"):
        pass
    w = TraceWindow("")
    w.step(0)
    w.close()


def test_async_checkpointer_latest_wins_and_text_jobs(tmp_path):
    from msrflute_b200.utils.async_ckpt import AsyncCheckpointer
    ck = AsyncCheckpointer(min_interval=0.05)
    flat = torch.arange(64.0)
    path = str(tmp_path / "latest.tar")
    for i in range(20):                                   # many submissions, few writes; the last one must win
        flat += 1
        ck.submit(path, {"model_state_dict": {"a": flat[:16].view(4, 4), "b": flat[16:]}, "round": i})
    ck.submit_text(str(tmp_path / "status.json"), json.dumps({"i": 20}))
    ck.flush()
    sd = torch.load(path, weights_only=False)
    assert sd["round"] == 19 and torch.equal(sd["model_state_dict"]["b"], flat[16:])
    assert json.load(open(tmp_path / "status.json")) == {"i": 20}
    assert ck.written < 22 and ck.coalesced > 0
    ck.close()
