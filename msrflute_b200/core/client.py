"""Simulated FL client (ref. ``core/client.py``).

``Client`` is a short-lived handle for (round, client ids).  The static methods carry the behaviour:

``get_train_dataset``  cache the whole training set per rank (ref :77-99)
``get_data``           slice per-user structures out of a dataset (ref :102-124)
``process_round``      local training + pseudo-gradient + strategy payload (+ personalization, privacy
                       metrics); returns the ``client_output`` dict ``{cs,tl,mg,vg,ng,rg,ns,pl,[ps],[wt],ts}``
                       (ref :227-511)
``run_testvalidate``   evaluate a chunk of val/test users (ref :127-222)

B200-first differences.  The reference rebuilds everything per client — ``deepcopy(config)`` twice, a fresh
optimizer, two ``deepcopy(optimizer.state_dict())``, a per-tensor ``p.data = data.clone().cuda()`` — and ships
the payload to the CPU.  Here a persistent per-model :class:`ClientContext` keeps the optimizer / trainer /
CUDA-graphed step alive across clients; receiving the global model is ONE arena copy; the pseudo-gradient
``w_global − w_local`` is ONE kernel over the arena; payloads stay on the device.
"""
from __future__ import annotations

import copy
import logging
import os
import time

import numpy as np
import torch

from ..extensions import privacy
from ..extensions.privacy import metrics as privacy_metrics
from ..models import make_model
from ..ops import arena_ops
from ..parallel.arena import adopt_module, module_arena
from ..utils import (ScheduledSamplingScheduler, alpha_update, convex_inference, make_optimizer, print_rank,
                     to_device)
from ..utils.dataloaders_utils import get_dataset, make_test_dataloader, make_train_dataloader, make_val_dataloader
from .config import ConfigNode
from .strategies import select_strategy
from .trainer import Trainer, run_validation_generic, set_component_wise_lr

# per-rank dataset cache (the reference uses module globals the same way, client.py:86-99)
train_dataset = None
trainset_unlab = None
trainset_unlab_rand = None


class ClientContext:
    """State that survives from one simulated client to the next on a worker."""

    def __init__(self, model):
        self.model = model
        self.optimizer = None
        self.optimizer_key = None
        self.trainer = None
        self.global_flat = None      # flat copy of the weights received this round
        self.global_key = None
        self.graphed = None          # core.graphed.GraphedTrainStep (optional)

    @staticmethod
    def of(model) -> "ClientContext":
        ctx = getattr(model, "_flute_client_ctx", None)
        if ctx is None:
            ctx = ClientContext(model)
            model._flute_client_ctx = ctx
        return ctx


def _install_parameters(model, model_parameters, send_dicts=False):
    """Copy the server's weights into the worker's persistent model.  Returns a flat tensor holding the
    received *parameters* (for the pseudo-gradient) when the model is arena-backed, else the list."""
    if send_dicts:
        sd = model.state_dict()
        keys = list(sd.keys())
        with torch.no_grad():
            for k, t in zip(keys, model_parameters):
                sd[k].copy_(t.to(sd[k].device))
        return None
    ar = module_arena(model)
    if ar is None and next(model.parameters(), None) is not None:
        try:
            ar = adopt_module(model, with_grad=True)
        except ValueError:
            ar = None
    if torch.is_tensor(model_parameters):                 # flat fast path
        flat = model_parameters
        if ar is not None and flat.numel() == ar[0].flat.numel():
            ar[0].flat.copy_(flat, non_blocking=True)
            return flat if flat.device == ar[0].flat.device else ar[0].flat.clone()
        model_parameters = ar[0].layout.views(flat) if ar is not None else [flat]
    with torch.no_grad():
        for p, t in zip(model.parameters(), model_parameters):
            p.data.copy_(t.to(p.device, non_blocking=True))
    if ar is not None:
        return ar[0].flat.clone()
    return [t.detach().clone() for t in model_parameters]


class Client:
    def __init__(self, client_id, config, send_gradients):
        self.client_id = client_id
        self.config = config              # read-only here; nothing below mutates the shared config
        self.send_gradients = send_gradients

    def get_client_data(self, dataset=None):
        return self.client_id, self.get_data(self.client_id, dataset), self.config, self.send_gradients

    # ------------------------------------------------------------------ data
    @staticmethod
    def get_train_dataset(data_path, client_train_config, task):
        """Load (once per rank) the training set(s); returns the number of users."""
        global train_dataset, trainset_unlab, trainset_unlab_rand
        train_dataset = get_dataset(data_path, client_train_config, task, mode="train")
        if task == "semisupervision":
            trainset_unlab = get_dataset(data_path, client_train_config, task, mode="train", user_idx=-2)
            trainset_unlab_rand = get_dataset(data_path, client_train_config, task, mode="train", user_idx=-3)
        else:
            trainset_unlab = trainset_unlab_rand = None
        return len(train_dataset.user_list)

    @staticmethod
    def get_data(clients, dataset):
        """Per-user slices ``{'users','num_samples','user_data'[,'user_data_label']}``; a list with one entry
        (three for semisupervision)."""
        if dataset is None:
            datasets = [train_dataset, trainset_unlab, trainset_unlab_rand] if trainset_unlab is not None \
                else [train_dataset]
        else:
            datasets = [dataset]
        has_labels = hasattr(datasets[0], "user_data_label")
        out = []
        for ds in datasets:
            users = [ds.user_list[c] for c in clients]
            st = {"users": users, "num_samples": [ds.num_samples[c] for c in clients],
                  "user_data": {u: ds.user_data[u] for u in users}}
            if has_labels:
                st["user_data_label"] = {u: ds.user_data_label[u] for u in users}
            node = ConfigNode()
            for k, v in st.items():           # shallow: do NOT wrap/copy the (large) per-user arrays
                dict.__setitem__(node, k, v)
            out.append(node)
        return out

    # ------------------------------------------------------------ evaluation
    @staticmethod
    def run_testvalidate(client_data, server_data, mode, model):
        _, data_strcts, config, _ = client_data
        _, model_parameters, iteration = server_data
        data_strct = data_strcts[0]
        data_config = config["server_config"]["data_config"][mode]
        want_logits = data_config.get("wantLogits", False)
        send_dicts = config["server_config"].get("send_dicts", False)
        task = config["server_config"]["task"]
        make_loader = make_test_dataloader if mode == "test" else make_val_dataloader
        dataloader = make_loader(data_config, data_path=None, task=task, data_strct=data_strct)
        model = to_device(model)
        if model_parameters is not None:
            _install_parameters(model, model_parameters, send_dicts)
        num_instances = sum(data_strct["num_samples"])
        output, metrics = run_validation_generic(model, dataloader)
        if config["server_config"].get("type", "model_optimization") == "personalization":
            Client._personalized_eval(config, data_strct, mode, data_config, task, output, metrics, make_loader)
        return (output if want_logits else None), metrics, num_instances

    @staticmethod
    def _personalized_eval(config, data_strct, mode, data_config, task, output, metrics, make_loader):
        """Interpolate global and per-user local model predictions (ref :190-219)."""
        model_path = config["model_path"]
        user = data_strct["users"][0]
        local_name = os.path.join(model_path, str(user) + "_model.tar")
        alpha_name = os.path.join(model_path, str(user) + "_alpha")
        if not (os.path.exists(local_name) and os.path.exists(alpha_name)):
            return
        local_model = make_model(config["model_config"])
        ckpt = torch.load(local_name, map_location=next(local_model.parameters()).device, weights_only=False)
        local_model.load_state_dict(ckpt["model_state_dict"])
        alpha = torch.load(alpha_name, weights_only=False)
        loader = make_loader(data_config, data_path=None, task=task, data_strct=data_strct)
        out_local, m_local = run_validation_generic(local_model, loader)
        metrics["loss"]["value"] = (metrics["loss"]["value"] + m_local["loss"]["value"]) / 2
        if isinstance(output, dict) and len(output.get("probabilities", [])):
            metrics["acc"]["value"] = convex_inference(output, out_local, alpha=alpha)

    # -------------------------------------------------------------- training
    @staticmethod
    def process_round(client_data, server_data, model, data_path, eps=1e-7):
        from . import federated
        if torch.cuda.is_available() and torch.cuda.device_count() == federated.size():
            torch.cuda.set_device(federated.local_rank())
        client_id, data_strcts, config, send_gradients = client_data
        initial_lr, model_parameters, iteration = server_data
        model_config, client_config = config["model_config"], config["client_config"]
        data_config = client_config["data_config"]["train"]
        task = client_config.get("task", {})
        trainer_config = client_config.get("trainer_config", {})
        privacy_metrics_config = config.get("privacy_metrics_config", None)
        strategy_algo = config["strategy"]
        strategy = select_strategy(strategy_algo)("client", config)
        send_dicts = config["server_config"].get("send_dicts", False)

        begin = time.time()
        client_stats = {}
        data_strct = data_strcts[0]
        user = data_strct["users"][0]
        print_rank("Loading : {}-th client with name: {}, {} samples".format(
            client_id[0], user, data_strct["num_samples"][0]), loglevel=logging.DEBUG)
        train_dataloader = make_train_dataloader(data_config, data_path, task=task, clientx=0, data_strct=data_strct)
        if model is None:
            model = make_model(model_config)
        model = to_device(model)
        ctx = ClientContext.of(model)

        # (1) receive the global model: one arena copy
        received = _install_parameters(model, model_parameters, send_dicts)

        # (2) optimizer: built once per worker, state cleared per client
        opt_cfg = dict(client_config["optimizer_config"])
        opt_cfg["lr"] = initial_lr
        opt_key = (tuple(sorted((k, str(v)) for k, v in opt_cfg.items() if k != "lr")),
                   tuple(trainer_config.get("updatable_names", [])))
        if ctx.optimizer is None or ctx.optimizer_key != opt_key:
            params = set_component_wise_lr(model, opt_cfg, trainer_config["updatable_names"]) \
                if "updatable_names" in trainer_config else model
            ctx.optimizer = make_optimizer(opt_cfg, params)
            ctx.optimizer_key = opt_key
            ctx.base_lrs = [g["lr"] for g in ctx.optimizer.param_groups]
        optimizer = ctx.optimizer
        optimizer.state.clear()
        for g, base in zip(optimizer.param_groups, ctx.base_lrs):
            g["lr"] = 0.0 if ("updatable_names" in trainer_config and base == 0.0) else initial_lr

        ss_scheduler = None
        if client_config.get("ss_config", None) is not None:
            ss_scheduler = ScheduledSamplingScheduler(model=model, **client_config["ss_config"])

        annealing_config = client_config.get("annealing_config", None)
        if ctx.trainer is None:
            ctx.trainer = Trainer(model=model, optimizer=optimizer, ss_scheduler=ss_scheduler,
                                  train_dataloader=train_dataloader, server_replay_config=client_config,
                                  max_grad_norm=data_config.get("max_grad_norm", None), anneal_config=None,
                                  num_skips_threshold=client_config.get("num_skips_threshold", -1),
                                  ignore_subtask=client_config["ignore_subtask"])
        trainer = ctx.trainer
        trainer.optimizer, trainer.ss_scheduler = optimizer, ss_scheduler
        trainer.train_dataloader = train_dataloader
        trainer.max_grad_norm = data_config.get("max_grad_norm", None)
        trainer.use_arena = True
        trainer.cached_batches = []
        trainer.step = 0
        trainer.lr_scheduler = None
        if annealing_config is not None:
            from ..utils import make_lr_scheduler
            for g in optimizer.param_groups:
                g.pop("initial_lr", None)
            trainer.lr_scheduler = make_lr_scheduler(annealing_config, optimizer)
        from . import graphed as _graphed
        _graphed.attach(ctx, trainer, client_config)

        assert "desired_max_samples" in data_config, "Missing 'desired_max_samples' entry in data config parameter"
        desired_max_samples = data_config["desired_max_samples"]
        end = time.time()
        client_stats["setup"] = end - begin
        begin_training = end

        # (3) local training
        trainer.model.train()
        apply_privacy_metrics = bool(privacy_metrics_config and privacy_metrics_config["apply_metrics"])
        algo_payload = None
        if strategy_algo == "FedLabels":
            datasets = [get_dataset(data_path, config, task, mode="train", test_only=False,
                                    data_strct=data_strcts[i], user_idx=0) for i in range(3)]
            algo_payload = {"strategy": "FedLabels", "data": datasets, "iter": iteration,
                            "config": client_config.get("semisupervision", None)}
        elif strategy_algo == "FedProx":
            algo_payload = {"strategy": "FedProx", "mu": client_config.get("mu", 0.001),
                            "reference_multiplicity": client_config.get("fedprox_reference_multiplicity", True)}
        train_loss, num_samples, algo_computation = trainer.train_desired_samples(
            desired_max_samples=desired_max_samples, apply_privacy_metrics=apply_privacy_metrics,
            algo_payload=algo_payload)
        print_rank("client={}: training loss={}".format(client_id[0], train_loss), loglevel=logging.DEBUG)
        assert "sum" in trainer.sufficient_stats and "mean" in trainer.sufficient_stats
        trainer.train_loss, trainer.num_samples, trainer.algo_computation = train_loss, num_samples, algo_computation

        # (4) pseudo-gradient  g = w_global − w_local   (one kernel on the arena)
        if not send_dicts:
            ar = module_arena(model)
            if ar is not None and torch.is_tensor(received) and ar[1] is not None:
                arena_ops.pseudo_grad(received, ar[0].flat, ar[1].flat)
            else:
                for p, data in zip(trainer.model.parameters(), received):
                    p.grad = data.to(p.device) - p.data
        payload = strategy.generate_client_payload(trainer) if send_gradients else None

        if config["server_config"].get("type", "model_optimization") == "personalization":
            Client._personalize(config, client_id, user, data_strct, data_config, data_path, task, trainer,
                                initial_lr, desired_max_samples, ss_scheduler)

        end = time.time()
        client_stats["training"] = end - begin_training
        client_stats["full cost"] = end - begin
        client_output = {
            "cs": client_stats, "tl": train_loss,
            "mg": trainer.sufficient_stats["mag"], "vg": trainer.sufficient_stats["var"],
            "ng": trainer.sufficient_stats["mean"], "rg": trainer.sufficient_stats["norm"],
            "ns": num_samples, "pl": payload,
        }
        if apply_privacy_metrics:
            Client._privacy_metrics(config, privacy_metrics_config, model_config, model, trainer, received,
                                    client_output)
        client_output["ts"] = time.time()
        return client_output

    # ------------------------------------------------------- personalization
    @staticmethod
    def _personalize(config, client_id, user, data_strct, data_config, data_path, task, trainer, initial_lr,
                     desired_max_samples, ss_scheduler):
        """Train the user's private model and update the mixing weight α (ref :387-443)."""
        client_config = config["client_config"]
        model_path = config["model_path"]
        alpha = client_config.get("convex_model_interp", 0.75)
        local_model = make_model(config["model_config"])
        loader = make_train_dataloader(data_config, data_path, task=task, clientx=0, data_strct=data_strct)
        opt_cfg = dict(client_config["optimizer_config"]); opt_cfg["lr"] = initial_lr
        local_trainer = Trainer(model=local_model, optimizer=make_optimizer(opt_cfg, local_model),
                                ss_scheduler=ss_scheduler, train_dataloader=loader,
                                server_replay_config=client_config,
                                max_grad_norm=data_config.get("max_grad_norm", None),
                                anneal_config=client_config.get("annealing_config", None),
                                ignore_subtask=client_config["ignore_subtask"], use_arena=False)
        local_name = os.path.join(model_path, str(user) + "_model.tar")
        alpha_name = os.path.join(model_path, str(user) + "_alpha")
        if os.path.exists(local_name):
            local_trainer.load(local_name, update_lr_scheduler=False, update_ss_scheduler=False)
        if os.path.exists(alpha_name):
            alpha = torch.load(alpha_name, weights_only=False)
        original = [p.detach().clone() for p in local_trainer.model.parameters()]
        local_trainer.model.train()
        loss, ns, _ = local_trainer.train_desired_samples(desired_max_samples=desired_max_samples,
                                                         apply_privacy_metrics=False)
        print_rank("client={}, user:{}: LOCAL training loss={}".format(client_id[0], user, loss), logging.DEBUG)
        local_trainer.save(model_path=model_path, config=None, token=str(user))
        for p, o in zip(local_trainer.model.parameters(), original):
            p.grad = o - p.data
        alpha = alpha_update(local_trainer.model, trainer.model, alpha, initial_lr)
        torch.save(alpha, alpha_name)

    # -------------------------------------------------------- privacy metrics
    @staticmethod
    def _privacy_metrics(config, pm_cfg, model_config, model, trainer, received, client_output):
        """Attack-style leakage metrics on this client's update; may drop the client (``wt = 0``)."""
        stats = {"Dropped clients": 0}
        batches, trainer.cached_batches = trainer.cached_batches, []
        gradients = privacy.unroll_network(model.named_parameters(), select_grad=True)[0]
        if pm_cfg.get("apply_indices_extraction", False):
            rank_cut = pm_cfg.get("allowed_word_rank", 9000)
            overlap, indices = privacy_metrics.extract_indices_from_embeddings(
                gradients, batches, model_config["embed_dim"], model_config["vocab_size"])
            max_overlap = pm_cfg.get("max_allowed_overlap", None)
            if max_overlap is not None and overlap > max_overlap:
                client_output["wt"] = 0.0
                stats["Dropped clients"] = 1
            stats["Extracted indices percentage"] = overlap
            stats["Words percentage above " + str(rank_cut) + " word rank"] = \
                float((indices > rank_cut).mean()) if len(indices) > 0 else 0
        if pm_cfg.get("apply_leakage_metric", False):
            ar = module_arena(model)
            if torch.is_tensor(received) and ar is not None:
                views = ar[0].layout.views(received)
            else:
                views = received
            orig = {n: t for (n, _), t in zip(model.named_parameters(), views)}
            leakage = privacy_metrics.practical_epsilon_leakage(
                orig, trainer.model, batches, pm_cfg["is_leakage_weighted"], np.exp(pm_cfg["max_leakage"]),
                pm_cfg["attacker_optimizer_config"])
            max_leak = pm_cfg.get("max_allowed_leakage", None)
            if max_leak is not None and leakage > max_leak:
                client_output["wt"] = 0.0
                stats["Dropped clients"] = 1
            stats["Practical epsilon (Max leakage)"] = leakage
        client_output["ps"] = stats
