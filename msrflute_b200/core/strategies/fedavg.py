"""Federated Averaging (ref. ``core/strategies/fedavg.py``).

Client: weight = number of samples processed, payload = weight·(w_global − w_local)
(``fedavg.py:80-89``), optional ``model_config.freeze_layer`` zeroing.  Server: drop
zero-weight payloads (``:109``), sum (in place when ``fast_aggregation``), divide by Σweight
(``:146-147``), optional cosine dump, ``update_model`` + lr scheduler (``:154-166``).
``strategy: FedProx`` maps here too; the proximal term lives in the trainer.
"""
import json
import logging
import os

import torch

from ...utils import compute_grad_cosines, print_rank
from .base import BaseStrategy
from .utils import aggregate_gradients_inplace, grad_arena, scale_gradients


class FedAvg(BaseStrategy):
    def __init__(self, mode, config, model_path=None):
        super().__init__(mode=mode, config=config, model_path=model_path)
        self.model_config = config["model_config"]
        self.client_config = config["client_config"]
        self.server_config = config["server_config"]
        self.dp_config = config.get("dp_config", None)
        if mode == "client":
            self.stats_on_smooth_grad = self.client_config.get("stats_on_smooth_grad", False)
        else:
            self.dump_norm_stats = self.config.get("dump_norm_stats", False)
            self.aggregate_fast = self.server_config.get("fast_aggregation", False)
            self.skip_model_update = False
            self.client_parameters_stack = []
            self.client_weights = []

    # ------------------------------------------------------------------ client
    def client_weight(self, trainer):
        return trainer.num_samples

    def _freeze(self, trainer):
        frozen = self.model_config.get("freeze_layer", None)
        if frozen:
            for n, p in trainer.model.named_parameters():
                if n == frozen and p.grad is not None:
                    print_rank("Setting gradient to zero for layer: {}".format(n), loglevel=logging.INFO)
                    p.grad.zero_()

    def generate_client_payload(self, trainer):
        self._require("client")
        if self.stats_on_smooth_grad:
            trainer.reset_gradient_power()
            trainer.estimate_sufficient_stats()
        weight = self.client_weight(trainer)
        scale_gradients(trainer.model, weight)
        self._freeze(trainer)
        return make_payload(trainer.model, weight)

    # ------------------------------------------------------------------ server
    @property
    def needs_individual_payloads(self):
        return self.mode == "server" and self.dump_norm_stats

    def process_individual_payload(self, worker_trainer, payload):
        self._require("server")
        if payload["weight"] == 0.0:
            return False
        self.client_weights.append(payload["weight"])
        if self.aggregate_fast:
            aggregate_gradients_inplace(worker_trainer.model, payload["gradients"], payload.get("flat"))
        else:
            self.client_parameters_stack.append(payload)
        return True

    def combine_payloads(self, worker_trainer, curr_iter, num_clients_curr_iter, total_clients, client_stats,
                         logger=None):
        self._require("server")
        cps_copy = None
        if self.dump_norm_stats:
            cps_copy = [[g.clone().detach() for g in x["gradients"]] for x in self.client_parameters_stack]
        weight_sum = self._aggregate_gradients(worker_trainer, num_clients_curr_iter, self.client_weights, logger)
        print_rank("Sum of weights: {}".format(weight_sum), loglevel=logging.DEBUG)
        if weight_sum:
            scale_gradients(worker_trainer.model, 1.0 / weight_sum)
        if cps_copy is not None:
            cosines = compute_grad_cosines(cps_copy, [p.grad.clone().detach() for p in worker_trainer.model.parameters()])
            with open(os.path.join(self.model_path, "cosines.txt"), "a", encoding="utf-8") as f:
                f.write("{}\n".format(json.dumps(cosines)))
        if self.skip_model_update is True:
            print_rank("Skipping model update")
            return
        worker_trainer.update_model()
        return worker_trainer.run_lr_scheduler(force_run_val=False)

    def _aggregate_gradients(self, worker_trainer, num_clients_curr_iter, client_weights, metric_logger=None):
        if not self.aggregate_fast:
            for payload in self.client_parameters_stack:
                aggregate_gradients_inplace(worker_trainer.model, payload["gradients"], payload.get("flat"))
        weight_sum = sum(client_weights)
        self.client_parameters_stack = []
        self.client_weights = []
        return weight_sum


def make_payload(model, weight):
    """``{'weight', 'gradients', ['flat']}``; gradients stay on the device as views of ONE flat buffer —
    the reference moves every tensor to the CPU here (``fedavg.py:89``)."""
    ga = grad_arena(model)
    if ga is not None:
        flat = ga.flat.detach().clone()
        pl = {"weight": weight, "gradients": ga.layout.views(flat), "flat": flat}
        wire = getattr(ga, "quant_wire", None)
        if wire is not None:                     # quantised payload (DGA): packed wire format between ranks
            pl["quant"] = wire
            ga.quant_wire = None
        return pl
    return {"weight": weight, "gradients": [p.grad.detach().clone() for p in model.parameters()]}
