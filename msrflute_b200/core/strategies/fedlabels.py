"""FedLabels — semi-supervised FL with separate supervised / unsupervised models
(ref. ``core/strategies/fedlabels.py``).

Each client returns TWO full state dicts: the model after supervised steps and the model after
pseudo-label training (``fedlabels.py:73-92``).  The server averages the supervised ones uniformly and
the unsupervised ones weighted by the number of pseudo-labels, and installs ``(sup + unsup)/2``
(``:145-149,188-216``).  Requires ``server_config.send_dicts: true`` semantics (state dicts, not
parameters, travel both ways).
"""
import logging

import numpy as np
import torch

from ...utils import print_rank
from .base import BaseStrategy


class FedLabels(BaseStrategy):
    def __init__(self, mode, config, model_path=None):
        super().__init__(mode=mode, config=config, model_path=model_path)
        self.model_config = config["model_config"]
        self.client_config = config["client_config"]
        self.server_config = config["server_config"]
        self.dp_config = config.get("dp_config", None)
        self.tmp_sup = None
        self.tmp_unsup = None
        if mode == "client":
            self.stats_on_smooth_grad = self.client_config.get("stats_on_smooth_grad", False)
        else:
            self.dump_norm_stats = self.config.get("dump_norm_stats", False)
            self.aggregate_fast = self.server_config.get("fast_aggregation", False)
            self.skip_model_update = False
            self.client_parameters_stack = []
            self.client_weights = []

    @property
    def needs_individual_payloads(self):
        return True

    def client_weight(self, trainer):
        return 1 if trainer.num_samples == 0 else trainer.num_samples

    def generate_client_payload(self, trainer):
        self._require("client")
        unsup = trainer.algo_computation
        if self.stats_on_smooth_grad:
            trainer.reset_gradient_power()
            trainer.estimate_sufficient_stats()
        sup = trainer.model.state_dict()
        sup_t = [sup[k].detach().clone() for k in sup]
        unsup_t = [unsup[k].detach().clone() for k in unsup]
        return {"weight": self.client_weight(trainer), "gradients": sup_t + unsup_t}

    def process_individual_payload(self, worker_trainer, payload):
        self._require("server")
        if payload["weight"] == 0.0:
            return False
        self.client_weights.append(payload["weight"])
        self.client_parameters_stack.append(payload["gradients"])
        return True

    def combine_payloads(self, worker_trainer, curr_iter, num_clients_curr_iter, total_clients, client_stats,
                         logger=None):
        self._require("server")
        if not self.client_parameters_stack:
            return None
        weight_sum, self.tmp_sup, self.tmp_unsup = self._aggregate(worker_trainer)
        print_rank("Sum of weights: {}".format(weight_sum), loglevel=logging.DEBUG)
        both = {}
        for k in self.tmp_unsup:
            s, u = self.tmp_sup[k], self.tmp_unsup[k]
            both[k] = ((s.float() + u.float()) / 2).to(s.dtype) if s.is_floating_point() else s
        worker_trainer.model.load_state_dict(both)
        if self.skip_model_update is True:
            print_rank("Skipping model update")
            return
        worker_trainer.optimizer.zero_grad(set_to_none=False)
        worker_trainer.update_model()
        return worker_trainer.run_lr_scheduler(force_run_val=False)

    def _aggregate(self, worker_trainer):
        keys = list(worker_trainer.model.state_dict().keys())
        dev = next(worker_trainer.model.parameters()).device
        half = len(self.client_parameters_stack[0]) // 2
        weights = np.asarray(self.client_weights, dtype=np.float64)
        weight_sum = float(weights.sum())
        ratio_sup = 1.0 / len(weights)
        ratio_unsup = weights / weight_sum

        def wavg(dicts, ratios):
            out = {}
            for ki, k in enumerate(keys):
                ref = dicts[0][ki]
                if not ref.is_floating_point():
                    out[k] = ref.to(dev).clone()
                    continue
                acc = torch.zeros_like(ref, dtype=torch.float32, device=dev)
                for d, r in zip(dicts, ratios):
                    acc.add_(d[ki].to(dev, torch.float32), alpha=float(r))
                out[k] = acc.to(ref.dtype)
            return out

        sup = wavg([c[:half] for c in self.client_parameters_stack], [ratio_sup] * len(weights))
        unsup = wavg([c[half:] for c in self.client_parameters_stack], ratio_unsup)
        self.client_parameters_stack, self.client_weights = [], []
        return weight_sum, sup, unsup
