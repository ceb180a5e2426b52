"""``OptimizationServer`` — the federated round loop (ref. ``core/server.py``).

Per round (ref ``train`` :215-528): sample clients → dispatch → consume client outputs through the strategy
→ ``combine_payloads`` (aggregate, DP, server optimizer) → optional server replay → evaluation cadence →
checkpoints (``latest`` every round, ``epoch<i>`` + best copies every ``model_backup_freq``) →
``status_log.json`` → metrics.  Output tree and file contents follow SURVEY §5.4 so reference tooling and
``resume_from_checkpoint`` keep working.

What is different underneath:

* the global model is a flat arena; what is dispatched each round is that buffer (no per-tensor CPU copies,
  ref :277-280);
* when the strategy does not need individual client gradients the round runs in *fused* mode: each rank
  accumulates ``Σ weight·(w_global − w_local)`` on its GPU, ranks are reduced once, and the server update is one
  fused pass (``ModelUpdater`` → ``ops.arena_ops.server_update``) — the reference's ``fast_aggregation``
  generalised across ranks;
* timing comes from CUDA events on the device (``utils/timing.py``) next to the reference's wall-clock stats.
"""
from __future__ import annotations

import cProfile
import json
import logging
import os
import pstats
import random
import shutil
import time
from collections import defaultdict

import numpy as np
import torch

from ..parallel.arena import adopt_module, module_arena
from ..utils import get_lr, print_rank, to_device, update_json_log
from ..utils.metrics_sink import get_run
from . import federated
from .engine import DeferredRound
from .evaluation import Evaluation
from .strategies import select_strategy
from . import trainer as _trainer_mod
from .trainer import ModelUpdater, Trainer, set_component_wise_lr

run = get_run()


class OptimizationServer(federated.Server):
    def __init__(self, num_clients, model, optimizer, ss_scheduler, data_path, model_path, server_train_dataloader,
                 config, idx_val_clients, idx_test_clients, single_worker=None, client_costs=None):
        super().__init__()
        self.client_idx_list = list(range(num_clients))
        self.client_costs = client_costs          # per-client num_samples → LPT load balancing
        self.config = config
        server_config = config["server_config"]
        decoder_config = config.get("decoder_config", None)
        self.max_iteration = server_config["max_iteration"]
        self.do_clustering = server_config.get("clustering", False)
        self.send_dicts = server_config.get("send_dicts", False)
        ncpi = server_config["num_clients_per_iteration"]
        self.num_clients_per_iteration = [int(x) for x in ncpi.split(",")] if isinstance(ncpi, str) else [ncpi]
        self.val_freq = server_config["val_freq"]
        self.req_freq = server_config["rec_freq"]
        self.evaluation = Evaluation(config, model_path, self.process_testvalidate, idx_val_clients,
                                     idx_test_clients, single_worker)
        self.metrics = dict()
        self.model_backup_freq = server_config.get("model_backup_freq", 100)
        self.worker_trainer_config = server_config.get("trainer_config", {})
        self.aggregate_median = server_config.get("aggregate_median", None)
        self.initial_lr_client = server_config.get("initial_lr_client", -1.0)
        self.lr_decay_factor = server_config.get("lr_decay_factor", 1.0)
        self.model_type = config["model_config"]["model_type"]
        self.quant_thresh = config["client_config"].get("quant_thresh", None)
        self.quant_bits = config["client_config"].get("quant_bits", 10)
        self.list_of_train_data = config["client_config"]["data_config"]["train"].get("list_of_train_data", None)
        self.data_path = data_path
        self.single_worker = single_worker

        max_grad_norm = server_config["data_config"]["train"].get("max_grad_norm", None) \
            if "train" in server_config["data_config"] else None
        model = to_device(model)
        if module_arena(model) is None:
            adopt_module(model, with_grad=True)
            if optimizer is not None:      # params kept identity; only their storage moved
                pass
        self.worker_trainer = ModelUpdater(model=model, optimizer=optimizer, ss_scheduler=ss_scheduler,
                                           train_dataloader=server_train_dataloader, val_dataloader=None,
                                           max_grad_norm=max_grad_norm,
                                           anneal_config=server_config["annealing_config"],
                                           model_type=self.model_type, decoder_config=decoder_config)
        self.metrics["worker_trainer"] = self.worker_trainer

        self.server_replay_iterations = None
        self.server_trainer = None
        if server_train_dataloader is not None:
            assert "server_replay_config" in server_config, "server_replay_config is not set"
            replay = server_config["server_replay_config"]
            assert "optimizer_config" in replay, "server-side replay training optimizer is not set"
            self.server_optimizer_config = replay["optimizer_config"]
            self.server_trainer_config = replay.get("trainer_config", {})
            self.server_replay_iterations = replay["server_iterations"]
            self.server_trainer = Trainer(
                model=model, optimizer=None, ss_scheduler=ss_scheduler, train_dataloader=server_train_dataloader,
                server_replay_config=replay,
                max_grad_norm=replay.get("max_grad_norm", server_config["data_config"]["train"].get("max_grad_norm", None)),
                anneal_config=replay.get("annealing_config", None), ignore_subtask=replay.get("ignore_subtask", False))

        self.skip_model_update = False
        self.train_loss = 0.0
        self.model_path = model_path
        self.best_model_criterion = server_config["best_model_criterion"]
        self.fall_back_to_best_model = server_config["fall_back_to_best_model"]
        self.last_model_path = os.path.join(self.model_path, "latest_model.tar")
        self.best_model_path = os.path.join(self.model_path, "best_val_{}_model.tar".format(self.best_model_criterion))
        self.log_path = os.path.join(self.model_path, "status_log.json")
        self.cur_iter_no = 0
        self.lr_weight = 1.0
        self.losses = []
        self.no_label_updates = 0
        if server_config.get("resume_from_checkpoint", False):
            self.load_saved_status()
        self.decoder_config = decoder_config
        self.spm_model = server_config["data_config"]["test"].get("spm_model", None)
        self.do_profiling = server_config.get("do_profiling", False)
        self.strategy = select_strategy(config["strategy"])("server", self.config, self.model_path)
        if hasattr(self.strategy, "run_validation"):
            self.strategy.run_validation = self._validation_for_strategy
        self.round_hooks = []          # callables(round_idx, metrics_payload) — used by bench.py for timing
        from . import trainer as _trainer_mod
        _trainer_mod.ASYNC_CHECKPOINTS["enabled"] = bool(
            server_config.get("b200", {}).get("async_checkpoint", torch.cuda.is_available()))
        print_rank(f"Server successfully instantiated strategy {self.strategy}", loglevel=logging.DEBUG)

    # ------------------------------------------------------------------ resume
    def load_saved_status(self):
        if os.path.exists(self.last_model_path):
            print_rank("Resuming from checkpoint model {}".format(self.last_model_path))
            self.worker_trainer.load(self.last_model_path, update_lr_scheduler=True, update_ss_scheduler=True)
            if self.server_trainer is not None:
                self.server_trainer.model = self.worker_trainer.model
        if os.path.exists(self.log_path):
            with open(self.log_path, "r") as f:
                e = json.load(f)
            self.cur_iter_no = e.get("i", 0)
            self.metrics["best_val_loss"] = e.get("best_val_loss", float("inf"))
            self.metrics["best_val_acc"] = e.get("best_val_acc", 0)
            self.metrics["best_test_loss"] = e.get("best_test_loss", float("inf"))
            self.metrics["best_test_acc"] = e.get("best_test_acc", 0)
            self.lr_weight = e.get("weight", 1.0)
            self.no_label_updates = e.get("num_label_updates", 0)
            print_rank(f"Resuming from status_log: cur_iter: {self.cur_iter_no}")

    def run(self):
        print_rank("server started")
        self.train()
        print_rank("server terminated")

    # -------------------------------------------------------------------- utils
    def _global_weights(self):
        """The buffer dispatched to workers: flat arena, or state-dict tensors when ``send_dicts``."""
        model = self.worker_trainer.model
        if self.send_dicts:
            sd = model.state_dict()
            return [sd[k].detach() for k in sd]
        ar = module_arena(model)
        return ar[0].flat if ar is not None else [p.data for p in model.parameters()]

    def _validation_for_strategy(self, mode="val"):
        m = self.evaluation.run_distributed_inference(mode, self.evaluation._global_values(self.worker_trainer))
        return (m.get("loss", {}).get("value"), m.get("acc", {}).get("value"))

    def _use_fused(self):
        if self.send_dicts or self.strategy.needs_individual_payloads:
            return False
        if self.config["server_config"].get("b200", {}).get("comm", "auto") == "p2p":
            return False
        if self.config["client_config"].get("type", "gradient_computation") != "optimization":
            return False
        return type(self.strategy).__name__ in ("FedAvg", "DGA")

    # -------------------------------------------------------------------- train
    def train(self):
        try:
            self.begin_training()
            self.run_rounds(self.max_iteration - self.cur_iter_no)
        finally:
            self.end_training()

    # The three phases of ``train`` are public so callers (bench.py, notebooks) can step rounds themselves.
    def begin_training(self):
        """Initial evaluation + the initial checkpoint dump (ref. ``server.py:232-255``)."""
        from ..utils.timing import RoundTimer, TraceWindow
        self._round_timer, self._trace = RoundTimer(), TraceWindow()
        self.run_stats = {k: [] for k in (
            "secsPerClientRound", "secsPerClient", "secsPerClientTraining", "secsPerClientSetup",
            "secsPerClientFull", "secsPerRoundHousekeeping", "secsPerRoundTotal", "communicationCosts")}
        run.log("Max iterations", self.max_iteration)
        self.worker_trainer.model = to_device(self.worker_trainer.model)
        eval_list = []
        if self.cur_iter_no == 0:
            if self.config["server_config"]["initial_rec"]:
                eval_list.append("test")
            if self.config["server_config"]["initial_val"]:
                eval_list.append("val")
                run.log("LR for agg. opt.", get_lr(self.worker_trainer.optimizer))
            print_rank("Running {} at itr={}".format(eval_list, self.cur_iter_no))
            if eval_list:
                self.metrics = self.evaluation.run(eval_list, self.metrics, metric_logger=run.log)
        print_rank("Saving Model Before Starting Training", loglevel=logging.INFO)
        for token in ["best_val_loss", "best_val_acc", "best_test_acc", "latest"]:
            self.worker_trainer.save(model_path=self.model_path, token=token, config=self.config["server_config"])
        self.worker_trainer.model.train()

    def run_rounds(self, n):
        """Run ``n`` federated rounds starting at ``self.cur_iter_no``; returns the last round's train loss."""
        last = None
        for i in range(self.cur_iter_no, min(self.cur_iter_no + n, self.max_iteration)):
            self._train_round(i, [])
            self.cur_iter_no = i + 1
            last = sum(self.train_loss) if self.train_loss else None
        return last

    def end_training(self):
        from ..utils.async_ckpt import flush_checkpoints
        if getattr(self, "_trace", None) is not None:
            self._trace.close()
        flush_checkpoints()
        self.terminate_workers(terminate=(not self.do_clustering))

    def _train_round(self, i, eval_list):
        begin = time.time()
        if getattr(self, "_trace", None) is not None:
            self._trace.step(i)
        metrics_payload = {}

        def log_metric(k, v):
            metrics_payload[k] = v

        print_rank("==== iteration {}".format(i))
        log_metric("Current iteration", i)
        initial_lr = self.initial_lr_client * self.lr_weight
        print_rank("Client learning rate {}".format(initial_lr), logging.DEBUG)
        ar = module_arena(self.worker_trainer.model)
        if ar is not None and ar[1] is not None:
            ar[1].zero_()                 # every .grad is a view of the flat gradient arena: one memset
        elif self.worker_trainer.optimizer is not None:
            self.worker_trainer.optimizer.zero_grad(set_to_none=False)
        else:
            self.worker_trainer.model.zero_grad()
        self.train_loss = []
        server_data = (initial_lr, self._global_weights(), i)

        if len(self.num_clients_per_iteration) > 1:
            num_clients_curr_iter = random.randint(self.num_clients_per_iteration[0], self.num_clients_per_iteration[1])
        else:
            num_clients_curr_iter = self.num_clients_per_iteration[0]
        log_metric("Clients for round", num_clients_curr_iter)

        extra = {}
        if self.quant_thresh is not None:
            self.config["client_config"]["quant_thresh"] *= self.config["client_config"].get("quant_anneal", 1.0)
            self.quant_thresh = self.config["client_config"]["quant_thresh"]
            log_metric("Quantization Thresh.", self.quant_thresh)
            extra["quant_thresh"] = self.quant_thresh

        sampled_idx_clients = random.sample(self.client_idx_list, num_clients_curr_iter) \
            if num_clients_curr_iter > 0 else self.client_idx_list
        costs = [self.client_costs[c] for c in sampled_idx_clients] if self.client_costs is not None else None

        clients_begin = time.time()
        client_mag_grads, client_mean_grads, client_var_grads, client_norm_grads = [], [], [], []
        for k in ("secsPerClient", "secsPerClientFull", "secsPerClientTraining", "secsPerClientSetup",
                  "communicationCosts"):
            self.run_stats[k].append([])

        pm_cfg = self.config.get("privacy_metrics_config", None)
        apply_privacy_metrics = bool(pm_cfg and pm_cfg["apply_metrics"])
        adaptive_leakage = apply_privacy_metrics and pm_cfg.get("adaptive_leakage_threshold", None)
        privacy_metrics_stats = defaultdict(list)
        if apply_privacy_metrics and pm_cfg.get("max_allowed_leakage", None) is not None:
            extra["max_allowed_leakage"] = pm_cfg["max_allowed_leakage"]

        profiler = None
        if self.do_profiling:
            profiler = cProfile.Profile()
            profiler.enable()

        fused = self._use_fused()
        fused_weights = []
        in_sync = fused and getattr(self, "_weights_in_sync", False)
        self._weights_in_sync = False
        sharded = self._sharded_request(fused, apply_privacy_metrics, i, num_clients_curr_iter)
        if sharded is not None:
            extra["sharded"] = sharded
        def consume(client_output):
            nonlocal num_clients_curr_iter
            client_stats = client_output["cs"]
            client_payload = client_output["pl"]
            if apply_privacy_metrics and "ps" in client_output:
                for metric, value in client_output["ps"].items():
                    privacy_metrics_stats[metric].append(value)
            self.run_stats["communicationCosts"][-1].append(time.time() - client_output["ts"])
            if client_output.get("wt", 1.0) == 0.0 and client_payload is not None:
                client_payload["weight"] = 0.0
            if fused:
                processed = client_payload is not None and client_payload["weight"] != 0.0
                if processed:
                    fused_weights.append(client_payload["weight"])
            else:
                processed = self.strategy.process_individual_payload(self.worker_trainer, client_payload)
            if not processed:
                print_rank("Dropping client", loglevel=logging.DEBUG)
                num_clients_curr_iter -= 1
                return
            self.train_loss.append(client_output["tl"])
            client_mag_grads.append(float(client_output["mg"]))
            client_mean_grads.append(float(client_output["ng"]))
            client_var_grads.append(float(client_output["vg"]))
            client_norm_grads.append(float(client_output["rg"]))
            client_end = time.time()
            self.run_stats["secsPerClientFull"][-1].append(client_stats["full cost"])
            self.run_stats["secsPerClientTraining"][-1].append(client_stats["training"])
            self.run_stats["secsPerClientSetup"][-1].append(client_stats["setup"])
            self.run_stats["secsPerClient"][-1].append(client_end - clients_begin)

        # Deferred read-back (device engine, single process): the clients' GPU work is only ENQUEUED here; the fused
        # server update, LR schedule and checkpoint snapshot are enqueued behind it and the per-client records are
        # read at the end of the round, so this round's host work overlaps its GPU work.
        deferred = None
        can_defer = fused and self._can_defer(apply_privacy_metrics)
        for client_output in self.process_clients(sampled_idx_clients, server_data, self.single_worker, costs=costs,
                                                  fused=fused, extra=extra, sync_weights=not in_sync, defer=can_defer):
            if isinstance(client_output, DeferredRound):
                deferred = client_output
                continue
            consume(client_output)

        fused_done = False
        if fused and deferred is not None and getattr(deferred, "sharded_done", False):
            fused_done = self._finish_sharded_round(deferred, i, num_clients_curr_iter, log_metric)
        elif fused and deferred is not None:
            fused_done = self._fused_server_update(None, i, num_clients_curr_iter, log_metric, wsum=deferred.weight_sum)
            if not fused_done:
                for o in deferred.resolve():
                    consume(o)
                deferred = None
                self._install_fused_aggregate(fused_weights)
        elif fused:
            fused_done = self._fused_server_update(fused_weights, i, num_clients_curr_iter, log_metric)
            if not fused_done:
                self._install_fused_aggregate(fused_weights)

        if self.do_profiling:
            profiler.disable()
            pstats.Stats(profiler).sort_stats("cumulative").print_stats(20)

        client_stats = (np.array(client_mag_grads), np.array(client_mean_grads), np.array(client_var_grads))
        if self.config.get("dump_norm_stats", False):
            with open(os.path.join(self.model_path, "norm_stats.txt"), "a", encoding="utf-8") as f:
                f.write("{}\n".format(json.dumps([float(x) for x in client_norm_grads])))
        if apply_privacy_metrics:
            for metric, values in privacy_metrics_stats.items():
                log_metric(metric, sum(values) if metric == "Dropped clients" else max(values))
        if type(adaptive_leakage) is float:
            values = privacy_metrics_stats["Practical epsilon (Max leakage)"]
            if values:
                new_threshold = sorted(values)[min(int(adaptive_leakage * len(values)), len(values) - 1)]
                print_rank("Updating leakage threshold to {}".format(new_threshold))
                self.config["privacy_metrics_config"]["max_allowed_leakage"] = new_threshold

        end = time.time()
        self.run_stats["secsPerClientRound"].append(end - begin)
        begin = end
        if deferred is None:
            log_metric("Training loss", sum(self.train_loss))

        if not fused_done:
            self.losses = self.strategy.combine_payloads(
                worker_trainer=self.worker_trainer, curr_iter=i, num_clients_curr_iter=num_clients_curr_iter,
                total_clients=len(self.client_idx_list), client_stats=client_stats, logger=log_metric)

        if self.server_trainer is not None:
            print_rank("Running replay iterations on server")
            if "updatable_names" in self.server_trainer_config:
                set_component_wise_lr(self.worker_trainer.model, self.server_optimizer_config,
                                      self.server_trainer_config["updatable_names"])
            self.server_trainer.prepare_iteration(self.worker_trainer.model)
            self.server_trainer.train_desired_samples(self.server_replay_iterations)
            if self.server_trainer.model is not self.worker_trainer.model:
                self.worker_trainer.model.load_state_dict(self.server_trainer.model.state_dict())

        self.worker_trainer.run_ss_scheduler()

        if ((i + 1) % self.val_freq) == 0:
            eval_list.append("val")
        if ((i + 1) % self.req_freq) == 0:
            eval_list.append("test")
        if len(eval_list) > 0:
            print_rank("Running {} at itr={}".format(eval_list, i + 1))
            self.metrics["worker_trainer"] = self.worker_trainer
            if getattr(self.strategy, "tmp_unsup", None) is not None:
                self.metrics["tmp_sup"] = self.strategy.tmp_sup
                self.metrics["tmp_unsup"] = self.strategy.tmp_unsup
            self.metrics = self.evaluation.run(eval_list, self.metrics, metric_logger=run.log)
            self.losses = self.evaluation.losses
            if "val" in eval_list:
                # client-lr decay on validation plateaus.  (In the reference this branch is dead code: the
                # list is cleared before it is tested, server.py:462-465; the documented behaviour is kept.)
                run.log("LR for agg. opt.", get_lr(self.worker_trainer.optimizer))
                best = self.metrics.get("best_val_loss", None)
                if self.losses[0] is not None and best is not None and not (self.losses[0] <= best):
                    self.lr_weight *= self.lr_decay_factor
                    print_rank("LOG: Client weight of learning rate {}..".format(self.lr_weight))

        self.backup_models(i)       # epoch<i>_* backups are queued behind pending writes (async_ckpt.submit_copy): no flush
        self.fall_back_to_prev_best_status()
        if deferred is not None:                       # everything of this round is enqueued: now read the records
            for o in deferred.resolve():
                consume(o)
            log_metric("Training loss", sum(self.train_loss))
        bg = _trainer_mod.ASYNC_CHECKPOINTS["enabled"]
        if len(self.metrics) > 1:
            update_json_log(self.log_path, background=bg, status_info={
                "i": i + 1,
                "best_val_loss": float(self.metrics.get("best_val_loss", float("inf"))),
                "best_val_acc": float(self.metrics.get("best_val_acc", 0)),
                "best_test_loss": float(self.metrics.get("best_test_loss", float("inf"))),
                "best_test_acc": float(self.metrics.get("best_test_acc", 0)),
                "weight": float(self.lr_weight),
                "num_label_updates": int(self.no_label_updates),
            })
        else:
            update_json_log(self.log_path, background=bg, status_info={
                "i": i + 1, "weight": float(self.lr_weight), "num_label_updates": int(self.no_label_updates)})
        end = time.time()
        self.run_stats["secsPerRoundHousekeeping"].append(end - begin)
        self.run_stats["secsPerRoundTotal"].append(
            self.run_stats["secsPerClientRound"][-1] + self.run_stats["secsPerRoundHousekeeping"][-1])
        log_metric("secsPerRoundTotal", self.run_stats["secsPerRoundTotal"][-1])
        if self.do_profiling:
            log_metric("secsPerClientRound", self.run_stats["secsPerClientRound"][-1])
            log_metric("secsPerRoundHousekeeping", self.run_stats["secsPerRoundHousekeeping"][-1])
            for metric in ("secsPerClient", "secsPerClientTraining", "secsPerClientFull", "secsPerClientSetup",
                           "communicationCosts"):
                vals = self.run_stats[metric][-1] or [0.0]
                log_metric(f"{metric}Mean", float(np.mean(vals)))
                log_metric(f"{metric}Median", float(np.median(vals)))
                log_metric(f"{metric}Max", float(max(vals)))
        if getattr(self, "_round_timer", None) is not None:      # device time of completed rounds (never blocks)
            self._round_timer.mark(i)
            for r, ms in self._round_timer.drain():
                run.log("devMsPerRound", ms)
        for k, v in metrics_payload.items():
            run.log(k, v)
        for hook in self.round_hooks:
            hook(i, metrics_payload)

    def _can_defer(self, apply_privacy_metrics):
        """Deferred read-back needs a round whose server side never looks at per-client host values."""
        dp = self.config.get("dp_config", None) or {}
        force = os.environ.get("FLUTE_FORCE_DEFER") == "1"        # CPU tests of the multi-rank deferred protocol
        return ((torch.cuda.is_available() or force) and not apply_privacy_metrics and not self.do_profiling
                and not dp.get("enable_global_dp", False) and not self.config.get("dump_norm_stats", False)
                and not self.strategy.skip_model_update)

    def _sharded_request(self, fused, apply_privacy_metrics, curr_iter, num_clients_curr_iter):
        """Transport v2 is used for a round when every rank can run its slice of the server step: symmetric-memory
        transport, deferred fused round, element-wise server optimizer, nothing that needs the aggregate on one GPU
        (norm logging / clipping, server replay, fall-back-to-best)."""
        comm = federated.get_comm()
        if not fused or comm.size == 1 or getattr(comm, "kind", "") != "symm" or not self._can_defer(apply_privacy_metrics):
            return None
        if not getattr(comm, "supports_sharded", lambda: False)():
            return None
        if self.server_trainer is not None or self.fall_back_to_best_model or self.config.get("dump_norm_stats", False):
            return None
        opt = self.worker_trainer.sharded_step_params()
        if opt is None:
            return None
        return {"opt": opt, "noise_scale": 0.0, "seed": 0}

    def _finish_sharded_round(self, deferred, curr_iter, num_clients_curr_iter, log_metric):
        worker = self.single_worker or federated._Runtime.worker
        self.worker_trainer.finish_sharded_step(worker.weight_buffer(), getattr(deferred, "state_mirrors", None))
        self._weights_in_sync = True
        self.strategy.client_weights, self.strategy.client_parameters_stack = [], []
        if type(self.strategy).__name__ == "DGA":
            from ..extensions import privacy
            privacy.update_privacy_accountant(self.config, len(self.client_idx_list), curr_iter=curr_iter,
                                              num_clients_curr_iter=num_clients_curr_iter, metric_logger=log_metric)
        self.losses = self.worker_trainer.run_lr_scheduler(force_run_val=False)
        return True

    def _fused_server_update(self, weights, curr_iter, num_clients_curr_iter, log_metric, wsum=None):
        """Fast path: the weighted pseudo-gradient sums (already reduced onto this rank, or peer-mapped when the
        transport is symmetric memory) go through ONE fused reduce/normalise/DP/optimizer/broadcast kernel
        (``ModelUpdater.fused_update``).  Returns False when some option needs the generic strategy path."""
        if not torch.cuda.is_available() or (wsum is None and not weights) or self.strategy.skip_model_update:
            return False
        if self.config.get("dump_norm_stats", False):
            return False
        worker = self.single_worker or federated._Runtime.worker
        comm = federated.get_comm()
        dp = self.config.get("dp_config", None) or {}
        noise_scale, seed, stats_out = 0.0, 0, None
        is_dga = type(self.strategy).__name__ == "DGA"
        if is_dga and dp.get("enable_global_dp", False):
            assert dp["enable_local_dp"], "global DP requires enable_local_dp (client-side clipping)"
            noise_scale = dp["global_sigma"] * dp["max_grad"] / num_clients_curr_iter
            from ..extensions.privacy import rng as dp_rng
            seed = dp_rng.dp_seed(stream=2)      # fresh OS entropy every round — never the reproducibility seed
            stats_out = torch.zeros(2, device=worker.accumulator().device)
        if wsum is None:
            wsum = torch.tensor(float(sum(weights)), device=worker.accumulator().device)
        accs = comm.peer_accumulators(worker.accumulator()) if hasattr(comm, "peer_accumulators") \
            else [worker.accumulator()]
        bcast = comm.peer_weight_buffers(worker.weight_buffer()) if hasattr(comm, "peer_weight_buffers") else None
        from ..utils.timing import PHASES
        with PHASES.phase("update_bcast"):
            ok = self.worker_trainer.fused_update(accs, wsum, noise_scale=noise_scale, seed=seed, bcast=bcast,
                                                  stats_out=stats_out)
            if ok and hasattr(comm, "round_done"):
                comm.round_done(worker.accumulator())
        if not ok:
            return False
        if bcast is not None and self.server_trainer is None and not self.fall_back_to_best_model:
            self._weights_in_sync = True
        self.strategy.client_weights, self.strategy.client_parameters_stack = [], []
        if stats_out is not None:
            log_metric("Gradient Norm", float(stats_out[0].item()))
        if is_dga:
            from ..extensions import privacy
            privacy.update_privacy_accountant(self.config, len(self.client_idx_list), curr_iter=curr_iter,
                                              num_clients_curr_iter=num_clients_curr_iter, metric_logger=log_metric)
        self.losses = self.worker_trainer.run_lr_scheduler(force_run_val=False)
        return True

    def _install_fused_aggregate(self, weights):
        """Move Σ_ranks Σ_clients weight·pseudo-grad (already reduced onto this rank) into ``model.grad`` and tell
        the strategy the gradients are aggregated in place (``fast_aggregation`` semantics)."""
        worker = self.single_worker or federated._Runtime.worker
        acc = worker.accumulator()
        ar = module_arena(self.worker_trainer.model)
        comm = federated.get_comm()
        if hasattr(comm, "peer_accumulators"):            # symmetric memory: peers were not reduced yet
            ar[1].flat.zero_()
            for a in comm.peer_accumulators(acc):
                ar[1].flat.add_(a)
            comm.round_done(acc)
        else:
            ar[1].flat.copy_(acc)
        acc.zero_()
        self.strategy.client_weights = list(weights)
        self.strategy.client_parameters_stack = []
        self.strategy.aggregate_fast = True

    # ----------------------------------------------------------------- backups
    def backup_models(self, i):
        self.worker_trainer.save(model_path=self.model_path, token="latest", config=self.config["server_config"])
        if (i % self.model_backup_freq) == 0:
            self.worker_trainer.save(model_path=self.model_path, token="epoch{}".format(i),
                                     config=self.config["server_config"])
            from . import trainer as _trainer_mod
            for body in ("best_val_acc", "best_val_loss", "best_test_acc"):
                src = os.path.join(self.model_path, "{}_model.tar".format(body))
                dst = os.path.join(self.model_path, "epoch{}_{}_model.tar".format(i, body))
                if _trainer_mod.ASYNC_CHECKPOINTS["enabled"]:
                    from ..utils.async_ckpt import get_checkpointer
                    get_checkpointer().submit_copy(src, dst)          # off the training thread, after pending writes
                elif os.path.exists(src):
                    shutil.copyfile(src, dst)

    def fall_back_to_prev_best_status(self):
        if not self.fall_back_to_best_model:
            return
        print_rank("falling back to model {}".format(self.best_model_path))
        tmp_lr = get_lr(self.worker_trainer.optimizer)
        self.worker_trainer.load(self.best_model_path, update_lr_scheduler=False, update_ss_scheduler=False)
        for g in self.worker_trainer.optimizer.param_groups:
            g["lr"] = tmp_lr
        if self.server_trainer is not None:
            self.server_trainer.model = self.worker_trainer.model


class PersonalizationServer(OptimizationServer):
    """Server for ``server_config.type: personalization`` (ref. ``experiments/cv/server.py`` — whose constructor
    lacks ``single_worker`` and therefore cannot be built by the reference's own entry point).  Orchestration is
    identical; the personalization logic lives in ``Client._personalize`` / ``Client._personalized_eval``."""


def select_server(server_type):
    return PersonalizationServer if server_type == "personalization" else OptimizationServer
