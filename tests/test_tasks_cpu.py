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


def test_privacy_attack_metrics_on_word_level_model(tmp_path):
    """``privacy_metrics_config.apply_metrics``: embedding-gradient index extraction, word-rank statistic and the
    practical-epsilon leakage attack run on the GRU LM (the reference's leakage metric cannot run on its own model:
    it expects a logit tensor its ``inference`` never returns) and their aggregates are logged per round."""
    out = str(tmp_path)
    with open(os.path.join(ROOT, "testing", "hello_world_nlg_gru.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["privacy_metrics_config"] = {
        "apply_metrics": True, "apply_indices_extraction": True, "allowed_word_rank": 100, "apply_leakage_metric": True,
        "max_leakage": 30.0, "adaptive_leakage_threshold": 0.95, "is_leakage_weighted": True,
        "attacker_optimizer_config": {"lr": 0.03, "type": "adamax", "amsgrad": False}}
    cfg_path = os.path.join(out, "config.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", FLUTE_ALLOW_FALLBACK="1")
    cmd = [sys.executable, os.path.join(ROOT, "e2e_trainer.py"), "-dataPath", out, "-outputPath", out, "-config", cfg_path,
           "-task", "nlg_gru", "-backend", "gloo", "-experiment", "pm"]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    keys = {json.loads(l)["k"] for l in open(os.path.join(out, "pm", "log", "metrics.jsonl"))}
    assert {"Practical epsilon (Max leakage)", "Extracted indices percentage", "Dropped clients"} <= keys
    assert any(k.startswith("Words percentage above") for k in keys)


def _run_cfg(task, cfg, out, name):
    cfg_path = os.path.join(out, "config.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", FLUTE_ALLOW_FALLBACK="1", FLUTE_SYNTH_USERS="12")
    cmd = [sys.executable, os.path.join(ROOT, "e2e_trainer.py"), "-dataPath", out, "-outputPath", out, "-config", cfg_path,
           "-task", task, "-backend", "gloo", "-experiment", name]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    return {json.loads(l)["k"] for l in open(os.path.join(out, name, "log", "metrics.jsonl"))}


def test_baseline_config4_shakespeare_with_quantized_gather(tmp_path):
    """BASELINE.json config #4 in miniature: LSTM char-LM, DGA with gradient quantization on the gather path."""
    with open(os.path.join(ROOT, "experiments", "nlp_rnn_fedshakespeare", "config.yaml")) as f:
        cfg = _shrink(yaml.safe_load(f))
    cfg["strategy"] = "DGA"
    cfg["server_config"]["aggregate_median"] = "mean"
    cfg["client_config"].update({"quant_thresh": 0.3, "quant_bits": 8, "quant_anneal": 0.95})
    keys = _run_cfg("nlp_rnn_fedshakespeare", cfg, str(tmp_path), "q")
    assert {"Quantization Thresh.", "Training loss"} <= keys


def test_baseline_config5_bert_mlm_with_global_dp(tmp_path):
    """BASELINE.json config #5 in miniature: BERT MLM, DGA aggregation with client clipping + server Gaussian noise and
    RDP accounting logged every round."""
    with open(os.path.join(ROOT, "testing", "hello_world_mlm_bert.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["strategy"] = "DGA"
    cfg["server_config"]["aggregate_median"] = "mean"
    cfg["dp_config"] = {"enable_local_dp": True, "enable_global_dp": True, "eps": -1.0, "max_grad": 1.0, "global_sigma": 0.01,
                        "max_weight": 1.0, "min_weight": 0.0, "delta": 1e-6, "weight_scaler": 1.0}
    keys = _run_cfg("mlm_bert", cfg, str(tmp_path), "dp")
    assert {"Gradient Norm", "dp_epsilon_rdp", "dp_sigma", "Training loss"} <= keys


def test_bert_masked_rows_head_equals_the_dense_hf_loss():
    """The MLM loss computed from the masked rows only (fixed-capacity selection, fused CE) equals HF's dense
    logits + CrossEntropyLoss(ignore_index=-100), and so do the gradients."""
    import torch
    from msrflute_b200.models.bert_mlm import BERT
    cfg = {"BERT": {"model": {"model_name": "tiny-bert", "tcgen05_attention": False},
                    "training": {"seed": 3, "batch_size": 4}}}
    m = BERT(cfg)
    m.train()
    for mod in m.modules():                      # dropout off: the two passes must see the same network
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    torch.manual_seed(0)
    ids = torch.randint(5, 900, (4, 32))
    labels = torch.full_like(ids, -100)
    pick = torch.rand(ids.shape) < 0.15
    labels[pick] = ids[pick]
    batch = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
    loss_sparse = m.compute_loss(dict(batch))
    loss_sparse.backward()
    g_sparse = [p.grad.clone() for p in m.parameters() if p.grad is not None]
    m.zero_grad()
    m.mlm_head_rows = 1.0
    loss_dense = m.compute_loss(dict(batch))
    loss_dense.backward()
    g_dense = [p.grad.clone() for p in m.parameters() if p.grad is not None]
    assert abs(float(loss_sparse) - float(loss_dense)) < 1e-5
    assert len(g_sparse) == len(g_dense)
    for a, b in zip(g_sparse, g_dense):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-4)
