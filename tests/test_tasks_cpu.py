"""Every shipped task runs end to end on CPU (single process, tiny synthetic data) through ``e2e_trainer.py``:
the four hello-world configs of ``testing/`` plus the remaining experiment folders with their own ``config.yaml``
cut down to two rounds."""
import json
import math
import os
import subprocess
import sys

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "testing"))
from test_e2e_trainer import run_pipeline  # noqa: E402

HELLO = ["nlg_gru", "ecg_cnn", "mlm_bert", "classif_cnn"]
SLOW = pytest.mark.skipif(os.environ.get("FLUTE_SLOW_TESTS") != "1",
                          reason="minutes of ResNet training on CPU; set FLUTE_SLOW_TESTS=1 (passes: 163 s / 292 s)")
OTHERS = ["cv_cnn_femnist", "nlp_rnn_fedshakespeare", "fednewsrec", pytest.param("cv", marks=SLOW),
          pytest.param("semisupervision", marks=SLOW)]


def _finite_training_loss(exp):
    losses = [json.loads(l)["v"] for l in open(os.path.join(exp, "log", "metrics.jsonl"))
              if json.loads(l)["k"] == "Training loss"]
    assert losses and all(math.isfinite(float(v)) for v in losses), losses


@pytest.mark.parametrize("task", HELLO)
def test_hello_world_single_process(task, tmp_path):
    out = str(tmp_path)
    rc = run_pipeline(task, out, nproc=1)
    assert rc == 0, open(os.path.join(out, "log_{}.txt".format(task))).read()[-3000:]
    _finite_training_loss(os.path.join(out, "hello"))


def _shrink(cfg):
    sc = cfg["server_config"]
    sc["max_iteration"] = 2
    sc["num_clients_per_iteration"] = 2
    sc["val_freq"], sc["rec_freq"] = 2, 100
    sc["initial_val"], sc["initial_rec"] = False, False
    for blk in (sc["data_config"].get("val"), sc["data_config"].get("test"), cfg["client_config"]["data_config"]["train"]):
        if blk:
            blk["num_workers"] = 0
    cfg["client_config"]["data_config"]["train"]["desired_max_samples"] = 40
    return cfg


@pytest.mark.parametrize("task", OTHERS)
def test_experiment_config_two_rounds(task, tmp_path):
    out = str(tmp_path)
    with open(os.path.join(ROOT, "experiments", task, "config.yaml")) as f:
        cfg = _shrink(yaml.safe_load(f))
    cfg_path = os.path.join(out, "config.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", FLUTE_ALLOW_FALLBACK="1", FLUTE_SYNTH_USERS="12")
    cmd = [sys.executable, os.path.join(ROOT, "e2e_trainer.py"), "-dataPath", out, "-outputPath", out, "-config", cfg_path,
           "-task", task, "-backend", "gloo", "-experiment", "smoke"]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    _finite_training_loss(os.path.join(out, "smoke"))
