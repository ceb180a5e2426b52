"""Hermetic end-to-end runs of the FL engine on CPU (gloo), the reference's own test strategy
(``testing/test_e2e_trainer.py``: launch e2e_trainer via torch.distributed.run, assert exit code 0) plus the
numerical / resume assertions the reference never makes."""
import json
import os
import subprocess
import sys

import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_data(tmp, users=12, per=24, seed=0):
    from msrflute_b200.data import synthetic
    tr = synthetic.make_vector_classification(users, per, 784, 10, seed=seed, fixed=True)
    te = synthetic.make_vector_classification(4, per, 784, 10, seed=seed + 1, fixed=True)
    torch.save(tr, os.path.join(tmp, "train.pt"))
    torch.save(te, os.path.join(tmp, "test.pt"))


def _config(tmp, rounds=3, resume=False, strategy="FedAvg", extra_server=None, extra_client=None, dp=None):
    with open(os.path.join(ROOT, "experiments", "cv_lr_mnist", "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    sc = cfg["server_config"]
    sc.update({"max_iteration": rounds, "num_clients_per_iteration": 4, "val_freq": 1, "rec_freq": 2,
               "initial_val": True, "resume_from_checkpoint": resume, "initial_lr_client": 0.5})
    sc["data_config"]["val"].update({"val_data": "test.pt", "batch_size": 64})
    sc["data_config"]["test"].update({"test_data": "test.pt", "batch_size": 64})
    sc.update(extra_server or {})
    cfg["client_config"]["data_config"]["train"].update({"list_of_train_data": "train.pt", "batch_size": 8})
    cfg["client_config"].update(extra_client or {})
    cfg["strategy"] = strategy
    if dp:
        cfg["dp_config"] = dp
    path = os.path.join(tmp, "cfg.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path


def _run(tmp, cfg_path, nproc=1, port=29611, timeout=900, extra_env=None):
    out = os.path.join(tmp, "out")
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
                "--master-port", str(port)]
    cmd += [os.path.join(ROOT, "e2e_trainer.py"), "-config", cfg_path, "-outputPath", out, "-dataPath", tmp,
            "-task", "cv_lr_mnist", "-backend", "gloo", "-experiment", "exp"]
    env = dict(os.environ, PYTHONPATH=ROOT, FLUTE_ALLOW_FALLBACK="1", CUDA_VISIBLE_DEVICES="")
    env.update(extra_env or {})
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return os.path.join(out, "exp"), r.stdout + r.stderr


def _metrics(exp):
    out = {}
    with open(os.path.join(exp, "log", "metrics.jsonl")) as f:
        for line in f:
            e = json.loads(line)
            out.setdefault(e["k"], []).append(e["v"])
    return out


def test_single_process_round_loop_output_tree_and_learning(tmp_path):
    tmp = str(tmp_path)
    _write_data(tmp)
    exp, log = _run(tmp, _config(tmp, rounds=6))
    models = os.path.join(exp, "models")
    for f in ("latest_model.tar", "best_val_loss_model.tar", "best_val_acc_model.tar", "best_test_acc_model.tar",
              "epoch0_model.tar", "config.yaml", "status_log.json"):
        assert os.path.exists(os.path.join(models, f)), f
    assert os.path.exists(os.path.join(exp, "FLUTE_config.yaml")) and os.path.exists(os.path.join(exp, "log", "log.out"))
    st = json.load(open(os.path.join(models, "status_log.json")))
    assert st["i"] == 6 and st["best_val_loss"] < float("inf")
    m = _metrics(exp)
    assert len(m["Training loss"]) == 6 and len(m["secsPerRoundTotal"]) == 6
    assert m["Val loss"][-1] < m["Val loss"][0]                          # the global model actually improves
    ck = torch.load(os.path.join(models, "latest_model.tar"), weights_only=False)
    assert set(ck) >= {"model_state_dict", "optimizer_state_dict", "lr_scheduler_state_dict"}
    assert ck["model_state_dict"]["net.linear.weight"].shape == (10, 784)


def test_resume_from_checkpoint_continues_at_saved_iteration(tmp_path):
    tmp = str(tmp_path)
    _write_data(tmp)
    exp, _ = _run(tmp, _config(tmp, rounds=2))
    w2 = torch.load(os.path.join(exp, "models", "latest_model.tar"), weights_only=False)["model_state_dict"]["net.linear.weight"].clone()
    exp, log = _run(tmp, _config(tmp, rounds=4, resume=True))
    assert "Resuming from status_log: cur_iter: 2" in log
    assert "==== iteration 2" in log and "==== iteration 0" not in log
    st = json.load(open(os.path.join(exp, "models", "status_log.json")))
    assert st["i"] == 4
    w4 = torch.load(os.path.join(exp, "models", "latest_model.tar"), weights_only=False)["model_state_dict"]["net.linear.weight"]
    assert not torch.equal(w2, w4)


@pytest.mark.parametrize("strategy,extra", [
    ("DGA", {"server": {"aggregate_median": "softmax", "stale_prob": 0.3}, "client": {"quant_thresh": 0.5, "quant_bits": 6, "quant_anneal": 0.9}}),
    ("DGA", {"server": {"aggregate_median": "mean", "fast_aggregation": True},
             "dp": {"enable_local_dp": True, "enable_global_dp": True, "eps": -1.0, "max_grad": 1.0, "global_sigma": 0.01,
                    "max_weight": 1.0, "min_weight": 0.0, "delta": 1e-6}}),
    ("FedProx", {"client": {"mu": 0.01}}),
    ("DGA", {"server": {"aggregate_median": "softmax", "wantRL": True,
                        "RL": {"initial_epsilon": 0.5, "final_epsilon": 0.1, "epsilon_gamma": 0.9, "max_replay_memory_size": 8,
                               "minibatch_size": 2, "network_params": "16,12,4", "marginal_update_RL": True,
                               "optimizer_config": {"type": "adam", "lr": 0.01},
                               "annealing_config": {"type": "step_lr", "step_interval": "epoch", "gamma": 1.0,
                                                    "step_size": 100}}}}),
])
def test_strategies_end_to_end(tmp_path, strategy, extra):
    tmp = str(tmp_path)
    _write_data(tmp)
    exp, log = _run(tmp, _config(tmp, rounds=3, strategy=strategy, extra_server=extra.get("server"),
                                 extra_client=extra.get("client"), dp=extra.get("dp")))
    m = _metrics(exp)
    assert len(m["Training loss"]) == 3
    if (extra.get("server") or {}).get("wantRL"):           # RL-learned aggregation weights: agent trained + saved
        assert "RL Running Loss" in m or "RL" in log
        assert any(f.startswith("rl_") for f in os.listdir(os.path.join(exp, "models")) + os.listdir(exp)) or "RL" in log
    if extra.get("dp"):
        assert "dp_epsilon_rdp" in m and "Gradient Norm" in m
    if "quant_thresh" in (extra.get("client") or {}):
        assert abs(m["Quantization Thresh."][-1] - 0.5 * 0.9 ** 3) < 1e-6
        assert "Stale Gradients Ratio" in m


def test_two_ranks_gloo_every_rank_trains_and_server_aggregates(tmp_path):
    """BASELINE config #1: LR-MNIST FedAvg, world_size=2 on CPU/gloo."""
    tmp = str(tmp_path)
    _write_data(tmp)
    exp, log = _run(tmp, _config(tmp, rounds=3), nproc=2)
    st = json.load(open(os.path.join(exp, "models", "status_log.json")))
    assert st["i"] == 3
    m = _metrics(exp)
    assert m["Val loss"][-1] < m["Val loss"][0]
    assert "Worker on node 1: process started" in log


def test_two_ranks_gloo_deferred_round_protocol(tmp_path):
    """The deferred multi-rank round (reduce Σw → reduce accumulators → records gathered at the end of the round) is
    what runs on multi-GPU jobs; ``FLUTE_FORCE_DEFER`` exercises the same control flow on CPU/gloo."""
    tmp = str(tmp_path)
    _write_data(tmp)
    exp, log = _run(tmp, _config(tmp, rounds=3), nproc=2, port=29617, extra_env={"FLUTE_FORCE_DEFER": "1"})
    st = json.load(open(os.path.join(exp, "models", "status_log.json")))
    assert st["i"] == 3
    m = _metrics(exp)
    assert m["Val loss"][-1] < m["Val loss"][0]
    assert len(m["Training loss"]) == 3


def test_dead_worker_is_detected_not_hung(tmp_path):
    """in the reference a dead worker hangs the job forever.  Here every control message / collective has
    a timeout and the launcher tears the job down: a worker killed by fault injection at round 1 makes the whole run
    exit non-zero within seconds."""
    import time
    tmp = str(tmp_path)
    _write_data(tmp)
    out = os.path.join(tmp, "out")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29619", os.path.join(ROOT, "e2e_trainer.py"), "-config",
           _config(tmp, rounds=50), "-outputPath", out, "-dataPath", tmp, "-task", "cv_lr_mnist", "-backend", "gloo",
           "-experiment", "exp"]
    env = dict(os.environ, PYTHONPATH=ROOT, FLUTE_ALLOW_FALLBACK="1", CUDA_VISIBLE_DEVICES="",
               FLUTE_FAULT_INJECT="1:1", FLUTE_COMM_TIMEOUT_S="20")
    t0 = time.time()
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert time.time() - t0 < 120
    assert "fault injection: rank 1 exits at round 1" in (r.stdout + r.stderr)


def test_norm_stats_cosines_fallback_and_server_replay(tmp_path):
    """Less-travelled reference options: ``dump_norm_stats`` (norm_stats.txt + cosines.txt), ``fall_back_to_best_model``
    and server-side replay training (``server_replay_config`` + ``server_config.data_config.train``)."""
    tmp = str(tmp_path)
    _write_data(tmp)
    cfgp = _config(tmp, rounds=3, extra_server={
        "fall_back_to_best_model": True,
        "server_replay_config": {"server_iterations": 2, "optimizer_config": {"type": "sgd", "lr": 0.05}}})
    with open(cfgp) as f:
        c = yaml.safe_load(f)
    c["dump_norm_stats"] = True
    c["server_config"]["data_config"]["train"] = {"batch_size": 16, "train_data_server": "test.pt",
                                                  "desired_max_samples": 64, "num_workers": 0}
    with open(cfgp, "w") as f:
        yaml.safe_dump(c, f)
    exp, log = _run(tmp, cfgp)
    files = os.listdir(os.path.join(exp, "models"))
    assert "norm_stats.txt" in files and "cosines.txt" in files
    cos = [json.loads(l) for l in open(os.path.join(exp, "models", "cosines.txt"))]
    assert len(cos) == 3 and all(-1.0001 <= v <= 1.0001 for row in cos for v in row)
    assert "Running replay iterations on server" in log and "falling back to model" in log
    assert len(_metrics(exp)["Training loss"]) == 3


@pytest.mark.parametrize("name,strategy,server,b200,nproc", [
    ("individual_payloads", "DGA", {"aggregate_median": "softmax", "stale_prob": 0.3}, None, 2),   # per-client flat payloads on the wire
    ("reference_layout", "FedAvg", {}, {"server_is_worker": False, "dispatch": "dynamic"}, 3),        # rank 0 = server only
    ("work_queue", "FedAvg", {}, {"dispatch": "work_queue", "device_engine": False}, 3),               # pull-based dispatch
    ("work_queue_individual", "DGA", {"aggregate_median": "softmax", "stale_prob": 0.3},
     {"dispatch": "work_queue", "server_is_worker": False}, 3),
])
def test_multi_rank_variants(tmp_path, name, strategy, server, b200, nproc):
    tmp = str(tmp_path)
    _write_data(tmp)
    cfgp = _config(tmp, rounds=3, strategy=strategy, extra_server=server)
    if b200:
        with open(cfgp) as f:
            c = yaml.safe_load(f)
        c["server_config"]["b200"] = b200
        with open(cfgp, "w") as f:
            yaml.safe_dump(c, f)
    exp, log = _run(tmp, cfgp, nproc=nproc, port=29731 + nproc)
    st = json.load(open(os.path.join(exp, "models", "status_log.json")))
    assert st["i"] == 3
    m = _metrics(exp)
    assert len(m["Training loss"]) == 3 and m["Val loss"][-1] < m["Val loss"][0]
    if name.startswith("work_queue"):
        import re
        # every client of every round was trained exactly once, by whichever rank pulled it
        per_round = {}
        for rank, k, n, rnd in re.findall(r"work queue: rank (\d+) trained (\d+) of (\d+) clients of round (\d+)", log):
            per_round.setdefault(int(rnd), []).append((int(k), int(n)))
        own = 0 if (b200 or {}).get("server_is_worker") is False else 1        # rank 0's own share is not logged by Worker.run
        assert len(per_round) == 3, log[-2000:]
        for rnd, parts in per_round.items():
            assert len(parts) == nproc - 1 and all(n == 4 for _, n in parts)
            assert sum(k for k, _ in parts) <= 4 and (own or sum(k for k, _ in parts) == 4)


def test_quantised_individual_payloads_travel_packed(tmp_path):
    """DGA with gradient quantisation on the individual-payload path (stale gradients force it), 2 ranks over gloo:
    worker -> server payloads are level codes + keep bitmap (ops.quant_ops.wire_encode), and the run produces the same
    model as shipping fp32 values (FLUTE_PACKED_WIRE=0)."""
    results = {}
    for mode in ("1", "0"):
        tmp = str(tmp_path / ("packed" + mode))
        os.makedirs(tmp)
        _write_data(tmp)
        cfgp = _config(tmp, rounds=2, strategy="DGA", extra_server={"aggregate_median": "softmax", "stale_prob": 0.3},
                       extra_client={"quant_thresh": 0.5, "quant_bits": 8, "quant_anneal": 1.0})
        exp, log = _run(tmp, cfgp, nproc=2, port=29761 + int(mode), extra_env={"FLUTE_PACKED_WIRE": mode})
        m = _metrics(exp)
        results[mode] = (m["Val loss"], log)
        assert json.load(open(os.path.join(exp, "models", "status_log.json")))["i"] == 2
    assert "packed messages" in results["1"][1] and "packed messages" not in results["0"][1]
    for a, b in zip(results["1"][0], results["0"][0]):
        assert abs(a - b) < 1e-4 * max(1.0, abs(b)), (results["1"][0], results["0"][0])
