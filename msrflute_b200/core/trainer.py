"""Local training runtime: ``ModelUpdater`` (server side) and ``Trainer`` (client / server replay).

Parity target ``core/trainer.py`` of the reference: ``TrainerBase`` :30-79,
``ModelUpdater`` :82-197 (clip → ``optimizer.step`` → ``zero_grad``; lr / ss
schedulers; save / load), ``Trainer`` :200-689 (``run_train_epoch`` :341-414,
FedProx :416-501, FedLabels semi-supervised :503-619, sufficient statistics
:263-312, ``reset_optimizer`` :640-651, ``prepare_iteration`` :624-638),
``run_validation_generic`` :690-723, ``set_component_wise_lr`` :725-751,
``save_model`` :753-775.

B200-first design:

* parameters/gradients live in a flat arena (``parallel/arena.py``); one
  mini-batch's post-backward work (global-norm clip, gradient sufficient
  statistics, SGD update, zero-grad) is ONE fused pass
  (``ops.arena_ops.fused_client_step``) instead of ``clip_grad_norm_`` + a
  per-parameter ``grad.clone().cpu().numpy()`` + ``optimizer.step()``;
* losses and statistics accumulate in device scalars — the host reads them
  once per client, not once per mini-batch (the reference calls
  ``loss.item()`` every step, ``trainer.py:397``);
* the whole forward/backward/update of a mini-batch can be captured in a CUDA
  graph (``core/graphed.py``) and replayed per step.
"""
from __future__ import annotations

import copy
import logging
import os
import random
import re

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ..ops import arena_ops
from ..parallel.arena import adopt_module, module_arena, rebind_grads
from ..utils import (get_lr, get_lr_all, make_lr_scheduler, make_optimizer, print_rank, to_device, torch_save,
                     try_except_save, write_yaml, get_label_VAT)
from .metrics import Metrics


def count_batch_samples(batch) -> int:
    """Samples (or tokens / frames) in a batch, as the reference counts them (``trainer.py:399-405``)."""
    if "attention_mask" in batch:
        return int((batch["attention_mask"] == 1).sum().item())
    if "total_frames" in batch:
        return int(batch["total_frames"])
    return len(batch["x"])


class TrainerBase:
    def __init__(self, model, train_dataloader, optimizer, max_grad_norm=None, ignore_subtask=True,
                 model_type="LanguageModel", decoder_config=None):
        self.model = model
        self.train_dataloader = train_dataloader
        self.optimizer = optimizer
        self.max_grad_norm = max_grad_norm
        self.model_type = model_type
        self.decoder_config = decoder_config
        self.step = 0
        self.ignore_subtask = ignore_subtask

    def epoch_boundary(self):
        return self.step % len(self.train_dataloader.create_loader()) == 0 and self.step != 0

    def train_desired_samples(self, desired_max_samples, apply_privacy_metrics):
        pass

    def save(self, model_path, token=None, config=None):
        save_model(model_path=model_path, config=config, model=self.model, optimizer=self.optimizer,
                   lr_scheduler=getattr(self, "lr_scheduler", None), ss_scheduler=getattr(self, "ss_scheduler", None),
                   token=token)

    def load(self, save_path, update_lr_scheduler, update_ss_scheduler):
        """Restore model/optimizer(/schedulers) from ``save_path`` if it exists."""
        from ..utils.async_ckpt import flush_checkpoints
        flush_checkpoints()
        if not os.path.isfile(save_path):
            return False
        print_rank("Loading checkpoint: {}".format(save_path))
        dev = next(self.model.parameters()).device
        ckpt = torch.load(save_path, map_location=dev, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        if self.optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        sd = ckpt.get("lr_scheduler_state_dict")
        if sd and getattr(self, "lr_scheduler", None) is not None and update_lr_scheduler:
            self.lr_scheduler.load_state_dict(sd)
        sd = ckpt.get("ss_scheduler_state_dict")
        if sd and getattr(self, "ss_scheduler", None) is not None and update_ss_scheduler:
            self.ss_scheduler.load_state_dict(sd)
        return True


class ModelUpdater(TrainerBase):
    """Applies an already-aggregated gradient to the global model (no data involved)."""

    def __init__(self, model, optimizer, ss_scheduler, train_dataloader, val_dataloader, max_grad_norm,
                 anneal_config, model_type="LanguageModel", decoder_config=None):
        super().__init__(model=model, train_dataloader=train_dataloader, optimizer=optimizer,
                         max_grad_norm=max_grad_norm, model_type=model_type, decoder_config=decoder_config)
        self.val_dataloader = val_dataloader
        self.annealing_type = anneal_config["type"] if anneal_config is not None else None
        self.lr_scheduler = make_lr_scheduler(anneal_config, self.optimizer)
        self.ss_scheduler = ss_scheduler

    def update_model(self):
        if self.max_grad_norm is not None:
            grad_norm = nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
            print_rank(f"clipped norm: {grad_norm} to {min(float(grad_norm), self.max_grad_norm)}", logging.DEBUG)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=False)

    # ---- fused device path ---------------------------------------------
    _FUSABLE = {"SGD": "sgd", "Adam": "adam", "AdamW": "adamW", "Adamax": "adamax", "LAMB": "lamb",
                "LarsSGD": "LarsSGD"}

    def fused_state(self):
        """Build (once) the arena-resident optimizer state used by ``ops.arena_ops.server_update``; the torch
        optimizer's ``state`` entries are re-pointed at views of those arenas so ``state_dict()`` checkpoints
        stay in the reference's format."""
        if getattr(self, "_fused", None) is not None:
            return self._fused
        opt = self.optimizer
        ar = module_arena(self.model)
        kind = self._FUSABLE.get(type(opt).__name__)
        from ..utils.optimizers import AdamW as _OurAdamW
        if ar is None or kind is None or len(opt.param_groups) != 1:
            return None
        if type(opt).__name__ == "AdamW" and not isinstance(opt, _OurAdamW):
            return None                  # torch.optim.AdamW has different decay semantics
        g = opt.param_groups[0]
        if g.get("amsgrad", False) or g.get("maximize", False) or not next(self.model.parameters()).is_cuda:
            return None
        w = ar[0]
        st = arena_ops.ServerOptState(
            kind, w.flat.numel(), w.flat.device, lr=g["lr"], betas=g.get("betas", (0.9, 0.999)),
            eps=g.get("eps", 1e-8), weight_decay=g.get("weight_decay", 0.0) or 0.0, momentum=g.get("momentum", 0.0) or 0.0,
            dampening=g.get("dampening", 0.0) or 0.0, nesterov=g.get("nesterov", False),
            correct_bias=g.get("correct_bias", True))
        self._fused = (st, w.layout.segments(w.flat.device))
        self._fused_kind = kind
        self._bind_fused_state()
        # Any later ``optimizer.load_state_dict`` (``ModelUpdater.load`` — called every round by
        # ``fall_back_to_prev_best_status`` — or the RL path) replaces the state entries with fresh tensors: import
        # them into the arenas and re-point the entries, otherwise the fused kernel would keep updating orphaned
        # buffers and checkpoints would carry frozen optimizer state.
        if not getattr(self, "_fused_hook", None):
            self._fused_hook = opt.register_load_state_dict_post_hook(lambda _o: self._bind_fused_state())
        return self._fused

    def _bind_fused_state(self):
        """Copy whatever per-parameter state the torch optimizer currently holds (fresh, resumed from a checkpoint
        or just re-loaded) into the arena-resident state and re-point ``optimizer.state`` at views of the arenas."""
        if getattr(self, "_fused", None) is None:
            return
        st, kind, opt = self._fused[0], self._fused_kind, self.optimizer
        lay = module_arena(self.model)[0].layout
        mv = lay.views(st.m) if st.m is not None else None
        vv = lay.views(st.v) if st.v is not None else None
        v_key = "exp_inf" if kind == "adamax" else "exp_avg_sq"
        step = None
        for i, p in enumerate(self.model.parameters()):
            s = opt.state[p]
            if kind in ("sgd", "LarsSGD"):
                if st.m is not None:
                    old = s.get("momentum_buffer")
                    if torch.is_tensor(old) and old.data_ptr() != mv[i].data_ptr():
                        mv[i].copy_(old)
                        st.step = max(st.step, 1)        # a loaded buffer is not a "first step" (buf = grad) any more
                    s["momentum_buffer"] = mv[i]
                continue
            for key, views in (("exp_avg", mv), (v_key, vv)):
                old = s.get(key)
                if torch.is_tensor(old) and old.data_ptr() != views[i].data_ptr():
                    views[i].copy_(old)
                s[key] = views[i]
            if "step" in s:
                step = int(s["step"].item()) if torch.is_tensor(s["step"]) else int(s["step"])
            else:
                s["step"] = 0 if kind in ("adamW", "lamb") else torch.tensor(0.0)
        if step is not None:
            st.step = step

    def sharded_step_params(self):
        """Hyper-parameters of the NEXT server step for the sharded (every-rank) update kernel, or ``None`` when this
        optimizer / configuration needs the single-GPU path (layer-wise trust ratios, server-side clipping)."""
        fs = self.fused_state()
        if fs is None or self.max_grad_norm:
            return None
        st = fs[0]
        if st.kind in ("lamb", "LarsSGD"):
            return None
        if (st.m is not None or st.v is not None) and st.step > 0 and not getattr(self, "_sharded_started", False):
            return None                  # resumed optimizer state lives on the server only: keep the single-GPU path
        self._sharded_started = True
        g = self.optimizer.param_groups[0]
        return {"code": st.code, "step": st.step + 1, "lr": float(g["lr"]), "betas": tuple(st.betas), "eps": st.eps,
                "weight_decay": st.weight_decay, "momentum": st.momentum, "dampening": st.dampening,
                "nesterov": st.nesterov, "correct_bias": st.correct_bias, "need_m": st.m is not None,
                "need_v": st.v is not None}

    def finish_sharded_step(self, w_buf, state_mirrors):
        """Server side of a sharded round: the new weights are in the (symmetric) weight buffer, the optimizer state of
        every slice was mirrored into ``state_mirrors`` — refresh the model arena and the checkpointable state."""
        st = self.fused_state()[0]
        module_arena(self.model)[0].flat.copy_(w_buf)
        if state_mirrors is not None:
            if st.m is not None and state_mirrors[0] is not None:
                st.m.copy_(state_mirrors[0])
            if st.v is not None and state_mirrors[1] is not None:
                st.v.copy_(state_mirrors[1])
        st.step += 1
        self.optimizer._opt_called = True
        for p in self.model.parameters():
            s = self.optimizer.state.get(p)
            if s is not None and "step" in s:
                s["step"] = st.step if not torch.is_tensor(s["step"]) else torch.tensor(float(st.step))

    def fused_update(self, accs, weight_sum, noise_scale=0.0, seed=0, bcast=None, stats_out=None, grad_out=None):
        """One fused pass: Σ_ranks acc / Σw (+noise) (+clip) → optimizer → broadcast.  Returns False when the
        configuration has no fused kernel (caller falls back to ``update_model``)."""
        fs = self.fused_state()
        if fs is None:
            return False
        st, segs = fs
        g = self.optimizer.param_groups[0]
        st.lr = float(g["lr"])
        w = module_arena(self.model)[0].flat
        arena_ops.server_update(w, accs, weight_sum, st, grad_out=grad_out, noise_scale=noise_scale, seed=seed,
                                max_grad_norm=self.max_grad_norm, segments=segs, bcast=bcast, zero_accs=True,
                                stats_out=stats_out)
        self.optimizer._opt_called = True          # the fused kernel WAS the optimizer step (lr_scheduler order check)
        for p in self.model.parameters():          # keep the per-parameter step counters in sync for checkpoints
            s = self.optimizer.state.get(p)
            if s is not None and "step" in s:
                s["step"] = st.step if not torch.is_tensor(s["step"]) else torch.tensor(float(st.step))
        return True

    def run_lr_scheduler(self, force_run_val=False):
        val_loss = val_acc = None
        if force_run_val is True or self.annealing_type == "val_loss":
            if self.val_dataloader is not None:
                _, metrics = run_validation_generic(self.model, self.val_dataloader)
                val_loss, val_acc = metrics["loss"]["value"], metrics["acc"]["value"]
        print_rank(f"LR all: {list(get_lr_all(self.optimizer))}", loglevel=logging.DEBUG)
        if self.lr_scheduler is not None:
            if self.annealing_type == "val_loss":
                if val_loss is not None:
                    self.lr_scheduler.step(val_loss)
            else:
                self.lr_scheduler.step()
        print_rank("LR AFTER lr_scheduler step: {}".format(get_lr(self.optimizer)), loglevel=logging.DEBUG)
        return (val_loss, val_acc)

    def run_ss_scheduler(self):
        if self.ss_scheduler is not None:
            self.ss_scheduler.step()


def _plain_sgd(opt) -> bool:
    return type(opt) is torch.optim.SGD and len(opt.param_groups) == 1 and not opt.param_groups[0].get("maximize", False)


def _fused_adamw(opt) -> bool:
    """The framework's own AdamW (HF / reference semantics) with one parameter group: runs as one fused arena kernel."""
    from ..utils.optimizers import AdamW
    return type(opt) is AdamW and len(opt.param_groups) == 1


class Trainer(TrainerBase):
    """Mini-batch SGD on one client's data (or on server replay data)."""

    def __init__(self, model, ss_scheduler, train_dataloader, server_replay_config=None, optimizer=None,
                 max_grad_norm=None, anneal_config=None, num_skips_threshold=-1, ignore_subtask=True,
                 use_arena=True):
        super().__init__(model=model, train_dataloader=train_dataloader, optimizer=optimizer,
                         max_grad_norm=max_grad_norm, ignore_subtask=ignore_subtask)
        self.server_replay_config = server_replay_config
        self.anneal_config = anneal_config
        self.lr_scheduler = None
        if self.optimizer is None and server_replay_config is not None and "optimizer" in server_replay_config:
            self.optimizer = make_optimizer(server_replay_config["optimizer_config"], model)
        if self.optimizer is not None and anneal_config is not None:
            self.lr_scheduler = make_lr_scheduler(anneal_config, self.optimizer)
        self.cached_batches = []
        self.ss_scheduler = ss_scheduler
        self.use_arena = use_arena
        self.step_fn = None          # optional CUDA-graphed replacement for ``_train_step`` (core/graphed.py)
        self.sufficient_stats = {}
        self._dev_state = None
        self.reset_gradient_power()

    # -- arena / device state -------------------------------------------------
    def _arena(self):
        if not self.use_arena:
            return None
        ar = module_arena(self.model)
        if ar is None or ar[1] is None:
            try:
                ar = adopt_module(self.model, with_grad=True)
            except ValueError:
                return None
        return ar

    def device_state(self):
        """(hyper[1,4], stats[1,4], loss_sum[1]) device tensors shared with graphed steps."""
        dev = next(self.model.parameters()).device
        if self._dev_state is None or self._dev_state[0].device != dev:
            self._dev_state = (arena_ops.make_hyper(1, dev), torch.zeros(1, 4, device=dev), torch.zeros(1, device=dev))
        return self._dev_state

    def _sync_hyper(self):
        hyper, _, _ = self.device_state()
        g = self.optimizer.param_groups[0] if self.optimizer is not None else {}
        vals = torch.tensor([[g.get("lr", 0.0), self.max_grad_norm or 0.0, g.get("weight_decay", 0.0),
                              g.get("momentum", 0.0)]], dtype=torch.float32)
        hyper.copy_(vals, non_blocking=True)

    # -- gradient statistics ---------------------------------------------------
    def reset_gradient_power(self):
        self.sum_grad = self.sum_grad2 = self.counter = 0
        if self._dev_state is not None:
            self._dev_state[1].zero_()

    def accumulate_gradient_power(self):
        """Σg, Σg², n over the current gradients — one flat reduction on the device."""
        ar = self._arena()
        if ar is not None:
            g = ar[1].flat
            s1, s2, n = g.sum(), (g * g).sum(), ar[1].layout.numel
        else:
            gs = [p.grad.detach().reshape(-1) for p in self.model.parameters() if p.grad is not None]
            flat = torch.cat(gs) if gs else torch.zeros(1)
            s1, s2, n = flat.sum(), (flat * flat).sum(), flat.numel()
        _, stats, _ = self.device_state()
        stats[0, arena_ops.S_SUM] += s1
        stats[0, arena_ops.S_SUMSQ] += s2
        stats[0, arena_ops.S_COUNT] += n
        return stats[0, 0], stats[0, 1], stats[0, 2]

    def estimate_sufficient_stats(self):
        self.accumulate_gradient_power()
        self._publish_stats()

    def _publish_stats(self):
        _, stats, _ = self.device_state()
        s = stats[0].detach().double().cpu().numpy()
        n = max(float(s[arena_ops.S_COUNT]), 1.0)
        self.sum_grad, self.sum_grad2, self.counter = float(s[0]), float(s[1]), n
        mean = np.float32(s[0] / n)
        mag = np.float32(np.sqrt(s[1] / n))
        self.sufficient_stats = {
            "n": n, "sum": np.float32(s[0]), "sq_sum": np.float32(s[1]),
            "var": np.float32(s[1] / n - float(mag) ** 2), "mean": mean, "mag": mag,
            "norm": np.float32(np.sqrt(s[1])),
        }

    # -- one mini-batch ---------------------------------------------------------
    def _loss(self, batch, apply_privacy_metrics=False):
        if self.ignore_subtask is True and hasattr(self.model, "single_task_loss"):
            return self.model.single_task_loss(batch)
        if apply_privacy_metrics:
            key = "x" if "x" in batch else ("input_ids" if "input_ids" in batch else None)
            if key is not None:
                self.cached_batches.append(to_device(batch[key]))
        return self.model.loss(batch)

    def _post_backward(self):
        """clip + stats + optimizer step (+ zero grad) after ``loss.backward()``."""
        ar = self._arena()
        hyper, stats, _ = self.device_state()
        if ar is not None:
            w, g = ar
            n = g.layout.numel
            if self.optimizer is None:
                arena_ops.clip_and_stats(g.flat, hyper, stats, n_logical=n)
            elif _plain_sgd(self.optimizer):
                grp = self.optimizer.param_groups[0]
                mom = None
                first = None
                if grp.get("momentum", 0) != 0:
                    mom, first = self._momentum_arena(w)
                arena_ops.fused_client_step(w.flat, g.flat, hyper, stats, mom, n_logical=n,
                                            nesterov=grp.get("nesterov", False), dampening=grp.get("dampening", 0.0),
                                            zero_grad=True, first_step=first)
                if first is not None:
                    first.zero_()
            elif _fused_adamw(self.optimizer):
                grp = self.optimizer.param_groups[0]
                m, v, step = self._adam_arena(w)
                step.add_(1)
                arena_ops.fused_client_adamw(w.flat, g.flat, m, v, step, hyper, stats, n_logical=n, betas=grp["betas"],
                                             eps=grp["eps"], correct_bias=grp.get("correct_bias", True), zero_grad=True)
            else:
                arena_ops.clip_and_stats(g.flat, hyper, stats, n_logical=n)
                self.optimizer.step()
                rebind_grads(self.model)
            return
        # generic (non-arena) path: same math with stock torch calls
        if self.max_grad_norm is not None:
            nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
        self.accumulate_gradient_power()
        if self.optimizer is not None:
            self.optimizer.step()

    def _adam_arena(self, w):
        """(m, v, step) flat AdamW state of the parameter arena; zeroed at the start of every client epoch (the client
        path clears ``optimizer.state`` per client, ref. ``core/client.py:343-347``)."""
        st = getattr(self, "_adam_state", None)
        if st is None or st[0].numel() != w.flat.numel() or st[0].device != w.flat.device:
            st = (torch.zeros_like(w.flat), torch.zeros_like(w.flat), torch.zeros(1, dtype=torch.int32, device=w.flat.device))
            self._adam_state = st
        return st

    def _momentum_arena(self, w):
        st = getattr(self, "_mom_state", None)
        if st is None or st[0].numel() != w.flat.numel() or st[0].device != w.flat.device:
            st = (torch.zeros_like(w.flat), torch.ones(1, dtype=torch.int32, device=w.flat.device))
            self._mom_state = st
        return st

    def _train_step(self, batch, apply_privacy_metrics=False, extra_loss=None):
        ar = self._arena()
        if ar is None and self.optimizer is not None:
            self.optimizer.zero_grad()
        loss = self._loss(batch, apply_privacy_metrics)
        if extra_loss is not None:
            loss = loss + extra_loss()
        loss.backward()
        self._post_backward()
        self.device_state()[2].add_(loss.detach().reshape(1).to(self.device_state()[2].dtype))
        return loss

    # -- epochs -------------------------------------------------------------------
    def train_desired_samples(self, desired_max_samples=None, apply_privacy_metrics=False, algo_payload=None):
        """Run one local epoch (capped at ``desired_max_samples``); returns (loss_sum, num_samples, algo_out)."""
        algo_computation = None
        if algo_payload is None:
            num_samples, train_loss = self.run_train_epoch(desired_max_samples, apply_privacy_metrics)
        elif algo_payload["strategy"] == "FedLabels":
            num_samples, train_loss, algo_computation = self.run_train_epoch_sup(
                desired_max_samples, apply_privacy_metrics, algo_payload)
        elif algo_payload["strategy"] == "FedProx":
            num_samples, train_loss = self.run_train_epoch_fedprox(desired_max_samples, apply_privacy_metrics, algo_payload)
        else:
            raise ValueError("unknown algo payload {}".format(algo_payload["strategy"]))
        return train_loss, num_samples, algo_computation

    def _begin_epoch(self):
        self.reset_gradient_power()
        ar = self._arena()
        if ar is not None:
            ar[1].zero_()
        else:
            self.model.zero_grad()
        self._sync_hyper()
        self.device_state()[2].zero_()
        if getattr(self, "_mom_state", None) is not None:
            self._mom_state[1].fill_(1)
        if getattr(self, "_adam_state", None) is not None:
            for t in self._adam_state:
                t.zero_()

    def _end_epoch(self):
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        loss_sum = float(self.device_state()[2].item())     # the ONE host sync of the epoch
        self._publish_stats()
        return loss_sum

    def run_train_epoch(self, desired_max_samples=None, apply_privacy_metrics=False, extra_loss=None):
        num_samples = 0
        self._begin_epoch()
        step = self.step_fn if (self.step_fn is not None and extra_loss is None and not apply_privacy_metrics) \
            else None
        for batch in self.train_dataloader.create_loader():
            if desired_max_samples is not None and num_samples >= desired_max_samples:
                break
            if step is not None:
                step(batch)
            else:
                self._train_step(batch, apply_privacy_metrics, extra_loss)
            num_samples += count_batch_samples(batch)
            self.step += 1
        return num_samples, self._end_epoch()

    def run_train_epoch_fedprox(self, desired_max_samples=None, apply_privacy_metrics=False, algo_payload=None):
        """FedProx: loss += Σ_p c_p·(μ/2)·‖w_p − w_global,p‖².

        The reference adds the *running* regulariser inside its parameter loop
        (``trainer.py:463-467``) so tensor i is counted (P−i) times; that behaviour is kept by default
        (``fedprox_reference_multiplicity``) and can be turned off to get textbook FedProx (c_p = 1)."""
        mu = algo_payload["mu"]
        ref_mult = algo_payload.get("reference_multiplicity", True)
        params = list(self.model.parameters())
        anchors = [p.detach().clone() for p in params]
        P = len(params)
        coefs = [(P - i) if ref_mult else 1 for i in range(P)]

        def prox():
            reg = 0.0
            for c, p, a in zip(coefs, params, anchors):
                reg = reg + c * (mu / 2) * (p - a).pow(2).sum()
            return reg

        return self.run_train_epoch(desired_max_samples, apply_privacy_metrics, extra_loss=prox)

    def run_train_epoch_sup(self, desired_max_samples=None, apply_privacy_metrics=False, algo_payload=None):
        """FedLabels semi-supervised local update (ref. ``trainer.py:503-619``).

        Supervised phase: ``train_ep`` SGD steps (lr 0.003, batch 64) on the labelled set with the received
        model.  Unsupervised phase (after ``burnout_round``): a copy of the *received* model is trained on
        pseudo-labels chosen by ``get_label_VAT`` from (local, server) predictions, with a KL consistency term
        and an L2 pull towards the received weights.  Returns
        ``(n_pseudo_labels, sup_loss/ensize, unsup_state_dict)``."""
        cfg = algo_payload["config"]
        round_ = algo_payload["iter"]
        sup_ds, unsup_ds, unsup_rand_ds = algo_payload["data"]
        self.reset_gradient_power()
        self.model.zero_grad()
        ce = torch.nn.CrossEntropyLoss()
        initial_net = copy.deepcopy(self.model)
        net_of = lambda m: getattr(m, "net", m)
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=0.003, momentum=0)
        num_samples, sum_loss = 0, 0.0
        for _ in range(int(cfg["train_ep"])):
            images, labels = next(iter(DataLoader(sup_ds, batch_size=64, shuffle=True)))
            self.model.zero_grad()
            labels = to_device(labels)
            loss = ce(net_of(self.model)(to_device(images)), labels)
            num_samples += len(labels)
            sum_loss += loss.item()
            loss.backward()
            self.optimizer.step()
        self.use_arena = False
        self.estimate_sufficient_stats()
        self.step += 1

        net = copy.deepcopy(initial_net)
        opt = torch.optim.SGD(net.parameters(), lr=cfg["eta"], momentum=0)
        total_est = 0
        if round_ >= cfg["burnout_round"]:
            kl = torch.nn.KLDivLoss(reduction="none", log_target=True)
            T = cfg["temp"]
            for _ in range(int(cfg["unsuptrain_ep"])):
                idx = random.sample(range(len(unsup_ds)), min(cfg["unl_bs"], len(unsup_ds)))
                images, _true = next(iter(DataLoader(torch.utils.data.Subset(unsup_ds, idx), batch_size=cfg["bs"])))
                images = to_device(images)
                initial_net.eval(); self.model.eval()
                with torch.no_grad():
                    out_local = net_of(initial_net)(images)
                    out_server = net_of(self.model)(images)
                p_local, p_server = torch.softmax(out_local / T, 1), torch.softmax(out_server / T, 1)
                est_labels, est_idx, est_var, _ = get_label_VAT(p_local, p_server, cfg["thre"], cfg["comp"])
                total_est += len(est_labels)
                if len(est_labels) == 0:
                    continue
                rand_images, _ = next(iter(DataLoader(torch.utils.data.Subset(unsup_rand_ds, idx), batch_size=cfg["bs"])))
                rand_images = to_device(rand_images)
                net.train()
                sel = torch.as_tensor(est_idx, device=images.device)
                out = net_of(net)(rand_images[sel] if cfg["uda"] == 1 else images[sel])
                out_norand = net_of(net)(images[sel])
                unsup_loss = ce(out, est_labels)
                kl_point = kl(torch.log_softmax(out_norand / T, 1), torch.log_softmax(out_server[sel] / T, 1))
                agree = (p_local[sel].argmax(1) == p_server[sel].argmax(1))
                consist = out.new_zeros(())
                if len(est_var) and bool(agree.any()):
                    consist = (kl_point.sum(1) * est_var)[agree].sum() / agree.sum()
                reg = sum(torch.nn.functional.mse_loss(p, q.detach()) for p, q in
                          zip(net.parameters(), initial_net.parameters()))
                opt.zero_grad()
                (cfg["unsup_lamb"] * unsup_loss + cfg["vat_consis"] * consist + cfg["l2_lambda"] * reg).backward()
                opt.step()
        self.model.train()
        return total_est, sum_loss / cfg["ensize"], net.state_dict()

    # -- misc --------------------------------------------------------------------
    def get_model(self):
        return copy.deepcopy(self.model)

    def prepare_iteration(self, model=None):
        """Sync with the global model and (re)build optimizer/scheduler before server replay."""
        if model is not None:
            if model is not self.model:
                self.model.load_state_dict(model.state_dict())
            self.lr_scheduler = None
            if self.optimizer is None and self.server_replay_config is not None and \
                    "optimizer_config" in self.server_replay_config:
                self.optimizer = make_optimizer(self.server_replay_config["optimizer_config"], self.model)
            if self.optimizer is not None and self.anneal_config is not None:
                self.lr_scheduler = make_lr_scheduler(self.anneal_config, self.optimizer)

    def reset_optimizer(self, optimizer_state_dict, annealing_config=None):
        assert self.optimizer is not None, "This trainer does not have an optimizer"
        self.optimizer.load_state_dict(optimizer_state_dict)
        self.lr_scheduler = None
        if annealing_config is not None:
            self.lr_scheduler = make_lr_scheduler(annealing_config, self.optimizer)


#: set by the server from ``server_config.b200.async_checkpoint`` (default: on when CUDA is available)
ASYNC_CHECKPOINTS = {"enabled": False}


def run_validation_generic(model, val_dataloader):
    """Evaluate ``model`` on a loader; returns ``(outputs, metrics)`` (ref. ``trainer.py:690-723``)."""
    model.set_eval()
    loader = val_dataloader.create_loader()
    return Metrics().compute_metrics(dataloader=loader, model=model)


def set_component_wise_lr(model, optimizer_config, updatable_names):
    """Param groups with lr 0 for every tensor whose name matches none of the regexes (ref. ``:725-751``)."""
    groups = []
    for name, p in model.named_parameters():
        hit = any(re.match(pat, name) is not None for pat in updatable_names)
        print_rank(("updating {} with lr = {}".format(name, optimizer_config["lr"])) if hit
                   else "freezing {}".format(name), logging.DEBUG)
        groups.append({"params": p, "lr": optimizer_config["lr"] if hit else 0.0})
    return groups


def save_model(model_path, config, model, optimizer, lr_scheduler, ss_scheduler, token=None):
    """Checkpoint writer; file layout identical to the reference (``trainer.py:753-775``)."""
    state = {
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
        "lr_scheduler_state_dict": lr_scheduler.state_dict() if lr_scheduler is not None else None,
    }
    if ss_scheduler is not None:
        state["ss_scheduler_state_dict"] = ss_scheduler.state_dict()
    save_path = os.path.join(model_path, "{}_model.tar".format(token) if token else "model.tar")
    print_rank("Saving model to: {}".format(save_path), logging.DEBUG)
    if ASYNC_CHECKPOINTS["enabled"]:
        from ..utils.async_ckpt import get_checkpointer
        get_checkpointer().submit(save_path, state)
    else:
        try_except_save(torch_save, state_or_model=state, save_path=save_path)
    if config is not None:
        try_except_save(write_yaml, config=config, save_path=os.path.join(model_path, "config.yaml"))
