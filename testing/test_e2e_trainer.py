"""End-to-end smoke of the four hello-world tasks — the reference's only functional test (``testing/
test_e2e_trainer.py``: launch ``e2e_trainer.py`` under ``torch.distributed.run --nproc_per_node=2`` per task and
assert exit code 0).  Same launcher, same four tasks, gloo on CPU when no GPU is visible; additionally checks that the
run produced the checkpoint / status files and finite metrics.

    python -m pytest testing/test_e2e_trainer.py -q            # from the repo root
"""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TASKS = ["nlg_gru", "ecg_cnn", "mlm_bert", "classif_cnn"]


def run_pipeline(task, out_dir, nproc=2, port=29640, timeout=900):
    config = os.path.join(ROOT, "testing", "hello_world_{}.yaml".format(task))
    backend = "nccl" if os.environ.get("FLUTE_TEST_BACKEND") == "nccl" else "gloo"
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                "--master-addr", "127.0.0.1", "--master-port", str(port)]
    cmd += [os.path.join(ROOT, "e2e_trainer.py"), "-dataPath", out_dir, "-outputPath", out_dir, "-config", config,
            "-task", task, "-backend", backend, "-experiment", "hello"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    if backend == "gloo":
        env.update(CUDA_VISIBLE_DEVICES="", FLUTE_ALLOW_FALLBACK="1")
    with open(os.path.join(out_dir, "log_{}.txt".format(task)), "w") as f:
        p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=f, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return p.returncode


@pytest.mark.parametrize("task", TASKS)
def test_hello_world(task, tmp_path):
    out = str(tmp_path)
    rc = run_pipeline(task, out, nproc=int(os.environ.get("FLUTE_TEST_NPROC", "2")), port=29640 + TASKS.index(task))
    log = open(os.path.join(out, "log_{}.txt".format(task))).read()
    assert rc == 0, log[-3000:]
    exp = os.path.join(out, "hello")
    assert os.path.exists(os.path.join(exp, "models", "latest_model.tar"))
    st = json.load(open(os.path.join(exp, "models", "status_log.json")))
    assert st["i"] >= 2
    losses = [json.loads(l)["v"] for l in open(os.path.join(exp, "log", "metrics.jsonl"))
              if json.loads(l)["k"] == "Training loss"]
    assert losses and all(math.isfinite(float(v)) for v in losses)
