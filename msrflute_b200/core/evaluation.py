"""Distributed evaluation (ref. ``core/evaluation.py``).

Validation/test users are packed into sample-balanced chunks (``make_eval_clients``, ref :185-216; one client
per user for personalization), the chunks are spread over the workers, every worker returns
``(output, metrics, count)`` per chunk and the server forms the sample-weighted mean of every metric
(ref :161-180), tracks ``best_<mode>_<metric>`` and writes ``best_val_<metric>_model.tar`` on improvement
(ref :92-109).  FedLabels additionally evaluates its separate supervised / unsupervised models (ref :76-86).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..utils import print_rank
from ..utils.metrics_sink import get_run
from . import federated
from .client import Client


class Evaluation:
    def __init__(self, config, model_path, process_testvalidate, idx_val_clients, idx_test_clients, single_worker):
        self.config = config
        self.model_path = model_path
        self.process_testvalidate = process_testvalidate
        self.server_type = config["server_config"].get("type", "model_optimization")
        self.idx_val_clients = idx_val_clients
        self.idx_test_clients = idx_test_clients
        self.send_dicts = config["server_config"].get("send_dicts", False)
        self.single_worker = single_worker
        self.losses = [None, None]
        self.metrics = {}
        self.logits = {}

    def _global_values(self, trainer):
        """What to ship for evaluation: the flat arena (fast path) or a state-dict tensor list (``send_dicts``)."""
        model = trainer.model
        if self.send_dicts:
            sd = model.state_dict()
            return [sd[k].detach() for k in sd]
        from ..parallel.arena import module_arena
        ar = module_arena(model)
        if ar is not None:
            return ar[0].flat
        return [p.data for p in model.parameters()]

    def run(self, eval_list, req, metric_logger=None):
        self.worker_trainer = req["worker_trainer"]
        values = self._global_values(self.worker_trainer)
        semi = "tmp_unsup" in req and req["tmp_unsup"] is not None
        if metric_logger is None:
            metric_logger = get_run().log
        for mode in eval_list:
            if self.config["server_config"].get("wantRL", False) and mode == "val":
                continue            # with RL the validation pass is driven by the strategy
            self.metrics = self.run_distributed_inference(mode, values)
            if not any(k.startswith("best_") for k in req):
                req = self.initialize_req(req)
            if semi:
                for tag, vals in (("Unsup", list(req["tmp_unsup"].values())), ("Sup", list(req["tmp_sup"].values()))):
                    m = self.run_distributed_inference(mode, vals, as_dict=True)
                    for key, value in m.items():
                        metric_logger(str(tag + mode + " " + key).capitalize(), value["value"])
            for key, value in self.metrics.items():
                metric_logger(str(mode + " " + key).capitalize(), value["value"])
                print_rank("LOG: {}_{}={}: best_{}_{}={}".format(
                    mode, key, value["value"], mode, key, req.get("best_" + mode + "_" + key)))
            for key, value in self.metrics.items():
                attr = "best_" + mode + "_" + key
                if attr not in req:
                    req[attr] = -1.0 if value["higher_is_better"] else float("inf")
                better = value["value"] > req[attr] if value["higher_is_better"] else value["value"] < req[attr]
                if better:
                    req[attr] = value["value"]
                    if mode == "val":
                        self.worker_trainer.save(model_path=self.model_path, token="best_" + mode + "_" + key,
                                                 config=self.config["server_config"])
        return req

    def initialize_req(self, req):
        for mode in ("test", "val"):
            for key, v in self.metrics.items():
                req["best_" + mode + "_" + key] = -1.0 if v["higher_is_better"] else float("inf")
        return req

    def run_distributed_inference(self, mode, model, as_dict=False):
        if mode == "val":
            clients = self.idx_val_clients
        elif mode == "test":
            clients = self.idx_test_clients
        else:
            raise NotImplementedError("Unsupported mode: {}".format(mode))
        return self.run_distributed_evaluation(mode, clients, model)

    def run_distributed_evaluation(self, mode, clients, model):
        total = 0
        sums, hib = {}, {}
        logits = {"predictions": [], "probabilities": [], "labels": []}
        server_data = (0.0, model, 0)
        for output, metrics, count in self.process_testvalidate(clients, server_data, mode, self.single_worker):
            for key, m in metrics.items():
                sums[key] = sums.get(key, 0.0) + m["value"] * count
                hib[key] = m["higher_is_better"]
            total += count
            if output is not None:
                for k in logits:
                    logits[k].append(output[k])
        if all(len(v) for v in logits.values()):
            logits = {k: np.concatenate(v) for k, v in logits.items()}
        self.logits = logits
        val_metrics = {k: {"value": s / max(total, 1), "higher_is_better": hib[k]} for k, s in sums.items()}
        self.losses = [val_metrics.get("loss", {}).get("value"), val_metrics.get("acc", {}).get("value")]
        return val_metrics


def make_eval_clients(dataset, config):
    """Yield evaluation ``Client`` chunks: users are packed until a chunk exceeds ``total/num_workers + 1``
    samples (one user per chunk for personalization)."""
    total = sum(dataset.num_samples)
    n_workers = federated.size() - 1 if federated.size() > 1 else federated.size()
    if config["server_config"].get("b200", {}).get("server_is_worker", True):
        n_workers = federated.size()
    threshold = total / max(n_workers, 1) + 1
    if config["server_config"].get("type", "model_optimization") == "personalization":
        for i in range(len(dataset.user_list)):
            yield Client([i], config, False)
        return
    cur, cur_total = [], 0
    for i in range(len(dataset.user_list)):
        cur.append(i)
        cur_total += dataset.num_samples[i]
        if cur_total > threshold:
            yield Client(cur, config, False)
            cur, cur_total = [], 0
    if cur:
        yield Client(cur, config, False)
