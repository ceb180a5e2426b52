"""Dynamic Gradient Aggregation (arXiv:2106.07578; ref. ``core/strategies/dga.py``).

Client weight (``dga.py:101-129``): 1.0 for ``aggregate_median: mean``; for ``softmax``
``exp(−β·x)`` with x = train_loss/num_samples | grad var | grad mean | grad magnitude chosen by
``weight_train_loss``, passed through ``filter_weight``.  Then local DP (``:133-134``), weighting +
``freeze_layer`` (``:142-146``), gradient quantization (``:149``).

Server (``:180-284``): optional RL-estimated weights, aggregation with *stale gradients* (each client's
update is deferred to the next round with probability ``stale_prob``), ``/Σw``, optional cosine dump,
global DP noise + RDP accounting (``:222-226``), model update, lr scheduler, RL training.
"""
import copy
import json
import logging
import math
import os

import numpy as np
import torch

from ...extensions import privacy
from ...extensions.quantization import quant_model
from ...utils import compute_grad_cosines, print_rank
from .base import BaseStrategy
from .fedavg import make_payload
from .utils import aggregate_gradients_inplace, filter_weight, scale_gradients

MIN_WEIGHT = 1e-7


class DGA(BaseStrategy):
    def __init__(self, mode, config, model_path=None):
        super().__init__(mode=mode, config=config, model_path=model_path)
        self.model_config = config["model_config"]
        self.client_config = config["client_config"]
        self.server_config = config["server_config"]
        self.dp_config = config.get("dp_config", None)
        if mode == "client":
            self.stats_on_smooth_grad = self.client_config.get("stats_on_smooth_grad", False)
            self.quant_threshold = self.client_config.get("quant_thresh", None)
            self.quant_bits = self.client_config.get("quant_bits", 10)
        else:
            self.dump_norm_stats = self.config.get("dump_norm_stats", False)
            self.aggregate_fast = self.server_config.get("fast_aggregation", False)
            self.want_rl = self.server_config.get("wantRL", False)
            self.stale_prob = self.server_config.get("stale_prob", 0.0)
            self.skip_model_update = False
            if self.aggregate_fast:
                if self.want_rl:
                    print_rank("RL is not possible with fast_aggregation; RL disabled", loglevel=logging.INFO)
                if self.stale_prob:
                    print_rank("stale gradients are not possible with fast_aggregation; stale_prob=0", logging.INFO)
                self.want_rl, self.stale_prob = False, 0.0
            self.rl = None
            if self.want_rl:
                from ...extensions.RL import RL
                self.rl = RL(config=self.server_config, model_path=self.model_path)
            self.client_parameters_stack = []
            self.client_parameters_stack_stale = []
            self.client_weights = []
            self.weight_sum_stale = 0.0
            self.losses = [None, None]
            self.run_validation = None      # set by the server: callable(mode) -> (loss, acc)

    # ------------------------------------------------------------------ client
    @property
    def uses_softmax(self):
        return self.server_config.get("aggregate_median", None) == "softmax"

    def training_signal(self, trainer):
        kind = self.server_config.get("weight_train_loss", "train_loss")
        if kind == "train_loss":
            return trainer.train_loss / max(trainer.num_samples, 1)
        if kind == "mag_var_loss":
            return trainer.sufficient_stats["var"]
        if kind == "mag_mean_loss":
            return trainer.sufficient_stats["mean"]
        return trainer.sufficient_stats["mag"]

    def client_weight(self, trainer):
        if not self.uses_softmax:
            return 1.0
        try:
            w = math.exp(-self.server_config["softmax_beta"] * float(self.training_signal(trainer)))
        except (OverflowError, ValueError):
            w = MIN_WEIGHT
        return filter_weight(w)

    def generate_client_payload(self, trainer):
        self._require("client")
        if self.stats_on_smooth_grad:
            trainer.reset_gradient_power()
            trainer.estimate_sufficient_stats()
        add_weight_noise = self.uses_softmax
        weight = self.client_weight(trainer)
        if weight > 0.0 and self.dp_config is not None and self.dp_config.get("enable_local_dp", False):
            weight = privacy.apply_local_dp(trainer, weight, self.dp_config, add_weight_noise)
        if not add_weight_noise:
            assert self.server_config.get("aggregate_median", "mean") == "mean"
            assert weight == 1.0
        scale_gradients(trainer.model, weight)
        frozen = self.model_config.get("freeze_layer", None)
        if frozen:
            for n, p in trainer.model.named_parameters():
                if n == frozen and p.grad is not None:
                    p.grad.zero_()
        quant_model(trainer.model, quant_threshold=self.quant_threshold, quant_bits=self.quant_bits,
                    global_stats=False)
        return make_payload(trainer.model, weight)

    # ------------------------------------------------------------------ server
    @property
    def needs_individual_payloads(self):
        return self.mode == "server" and (self.dump_norm_stats or self.want_rl or self.stale_prob > 0)

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
        rl_model = None
        if self.want_rl:
            rl_model = self._run_rl_inference(worker_trainer, self.client_weights, *client_stats)
        cps_copy = None
        if self.dump_norm_stats:
            cps_copy = [[g.clone().detach() for g in x["gradients"]] for x in self.client_parameters_stack]
        weights_this_round = list(self.client_weights)
        weight_sum = self._aggregate_gradients(worker_trainer, num_clients_curr_iter, self.client_weights, logger)
        print_rank("Sum of weights: {}".format(weight_sum), loglevel=logging.DEBUG)
        if weight_sum:
            scale_gradients(worker_trainer.model, 1.0 / weight_sum)
        if cps_copy is not None:
            cosines = compute_grad_cosines(cps_copy, [p.grad.clone().detach() for p in worker_trainer.model.parameters()])
            with open(os.path.join(self.model_path, "cosines.txt"), "a", encoding="utf-8") as f:
                f.write("{}\n".format(json.dumps(cosines)))
        privacy.apply_global_dp(self.config, worker_trainer.model, num_clients_curr_iter=num_clients_curr_iter,
                                select_grad=True, metric_logger=logger)
        eps = privacy.update_privacy_accountant(self.config, total_clients, curr_iter=curr_iter,
                                                num_clients_curr_iter=num_clients_curr_iter, metric_logger=logger)
        if eps:
            print_rank(f"DP result: {eps}")
        if self.skip_model_update is True:
            print_rank("Skipping model update")
            return
        worker_trainer.update_model()
        losses = worker_trainer.run_lr_scheduler(force_run_val=False)
        if self.want_rl:
            self.losses = list(losses) if losses is not None else [None, None]
            self._run_rl_training(worker_trainer, curr_iter, rl_model, weights_this_round, *client_stats, logger)
        return losses

    def _aggregate_gradients(self, worker_trainer, num_clients_curr_iter, client_weights, metric_logger=None):
        weight_sum = 0
        if not self.aggregate_fast:
            if metric_logger is not None:
                metric_logger("Stale Gradients Ratio",
                              len(self.client_parameters_stack_stale) / max(num_clients_curr_iter, 1))
            if self.client_parameters_stack_stale:
                weight_sum = self.weight_sum_stale
                for payload in self.client_parameters_stack_stale:
                    aggregate_gradients_inplace(worker_trainer.model, payload["gradients"], payload.get("flat"))
                self.client_parameters_stack_stale = []
                self.weight_sum_stale = 0
            for w, payload in zip(client_weights, self.client_parameters_stack):
                if np.random.random() > self.stale_prob:
                    aggregate_gradients_inplace(worker_trainer.model, payload["gradients"], payload.get("flat"))
                else:
                    self.weight_sum_stale += w
                    self.client_parameters_stack_stale.append(payload)
        weight_sum += sum(client_weights) - self.weight_sum_stale
        self.client_parameters_stack = []
        self.client_weights = []
        return weight_sum

    # ---------------------------------------------------------------------- RL
    # The reference's RL hooks reference attributes that do not exist on the strategy
    # (``self.worker_trainer`` at dga.py:300, ``self.run_distributed_inference`` at :367) and therefore
    # cannot run; the functions below implement the evident intent with the trainer passed in explicitly.
    def _rl_state(self, client_weights, mag, mean, var):
        return np.concatenate((np.asarray(client_weights, dtype=np.float64), mag, mean, var), axis=0)

    def _run_rl_inference(self, worker_trainer, client_weights, client_mag_grads, client_mean_grads,
                          client_var_grads):
        model = worker_trainer.model
        original = [p.data.detach().clone() for p in model.parameters()]
        opt_state = copy.deepcopy(worker_trainer.optimizer.state_dict())
        print_rank("RL estimation of the aggregation weights", loglevel=logging.INFO)
        state = self._rl_state(client_weights, client_mag_grads, client_mean_grads, client_var_grads)
        rl_weights = self.rl.forward(state).detach().cpu().numpy()
        if rl_weights.ndim > 1:
            rl_weights = rl_weights[-1, :]
        rl_weights = np.exp(rl_weights)[:len(client_weights)]
        rl_weights[~np.isfinite(rl_weights)] = 0
        worker_trainer.optimizer.zero_grad(set_to_none=False)
        weight_sum = 0.0
        for payload, orig_w, rl_w in zip(self.client_parameters_stack, client_weights, rl_weights):
            for p, g in zip(model.parameters(), payload["gradients"]):
                contrib = g.to(p.device) * (rl_w / orig_w)
                p.grad = contrib if p.grad is None else p.grad.add_(contrib)
            weight_sum += rl_w
        if weight_sum > 0:
            scale_gradients(model, 1.0 / weight_sum)
        worker_trainer.update_model()
        rl_losses = self.run_validation("val") if self.run_validation is not None else (None, None)
        rl_model = [p.data.detach().clone() for p in model.parameters()]
        for p, o in zip(model.parameters(), original):
            p.data.copy_(o)
        worker_trainer.optimizer.load_state_dict(opt_state)
        worker_trainer.optimizer.zero_grad(set_to_none=False)
        self.rl.set_weights(rl_weights)
        self.rl.set_losses(rl_losses)
        return rl_model

    def _run_rl_training(self, worker_trainer, iter, rl_model, client_weights, client_mag_grads, client_mean_grads,
                         client_var_grads, metric_logger):
        if None in self.losses and self.run_validation is not None:
            self.losses = list(self.run_validation("val"))
        if None in self.losses or None in self.rl.rl_losses:
            return
        print_rank("Performing RL training on the aggregation weights")
        gap = self.losses[1] - self.rl.rl_losses[1]       # baseline error − RL error  (index 1 = acc slot)
        adopt = False
        if abs(gap) < 0.001:
            reward = 0.1
            adopt = bool(self.server_config.get("marginal_update_RL", False))
        elif gap > 0:
            reward, adopt = 1.0, True
        else:
            reward = -1.0
        print_rank("Iter:{}  val_ACC={}  rl_val_ACC={}  reward={}".format(iter, self.losses[1], self.rl.rl_losses[1], reward))
        if adopt:
            self.losses = list(self.rl.rl_losses)
            for p, q in zip(worker_trainer.model.parameters(), rl_model):
                p.data.copy_(q)
        batch = (self._rl_state(client_weights, client_mag_grads, client_mean_grads, client_var_grads),
                 self.rl.rl_weights, [reward])
        self.rl.train(batch)
        self.rl.save(iter)
        if metric_logger is not None:
            metric_logger("RL Running Loss", self.rl.runningLoss)
            metric_logger("RL Rewards", reward)
