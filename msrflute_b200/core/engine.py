"""Device-resident multi-client training engine.

The reference trains one simulated client at a time per GPU, rebuilding python objects per client and copying every
gradient to the host every mini-batch.  A B200 is idle >95 % of the time under that regime: a ResNet-18 step on a
20×3×32×32 batch is ~2 GFLOP.  This engine keeps S *client slots* alive on the GPU and runs a whole wave of S clients
at once:

* **slots** — S model replicas whose parameters / gradients are rows of two ``[S, P]`` arenas; receiving the global
  model is one broadcast copy into ``W[S, P]``.
* **HBM-resident shards** — the whole federated training set is packed into one device tensor
  (``dataset.device_tensors``); a mini-batch for ALL slots is one on-device gather.  ``device_resident_data: false``
  keeps the pack in pinned host memory and copies only the sampled users' shards host→device each round.
* **wave-batched step** — all S clients advance one mini-batch together, captured in ONE CUDA graph (a 10-client ×
  5-step round of ResNet-18 is 5 graph replays).  Models with a slot-batched executor (``models/slot_resnet.py``) run
  every layer of all clients in one launch of this repo's kernels (tcgen05 convolutions addressing the per-client
  weights inside the arena, GroupNorm, max-pool, softmax-CE); other models go through
  ``torch.func.vmap(grad_and_value(loss))`` over the slot dimension.  ONE fused clip/statistics/SGD kernel pair over
  ``[S, P]`` ends the step.
* **compact slot arenas** — parameters whose gradient is structurally zero (filter taps that only see padding) are
  not replicated per client (``SlotBatchedResNet.plan_compact``): 4.4 M instead of 11.7 M elements per slot for
  ResNet-18 on 32×32 inputs; broadcast / gather go through an index map.
* **per-slot path** — when slots cannot move in lock-step (ragged last batches, models with buffers, ops without a
  vmap rule) each slot replays its own captured step on its own stream.
* **no host syncs** — losses, gradient statistics and aggregation weights stay in device tensors; the host reads one
  small ``[clients, 8]`` table per round, and with ``defer=True`` only after the caller has enqueued the server side of
  the round (:class:`DeferredRound`).
* **fused gather** — at the end of a wave ONE kernel does ``acc += Σ_s weight_s·(w_global − w_s)`` over all slots
  (``ops.arena_ops.accumulate_pseudo_grad``); per-client pseudo-gradients are never materialised.

It produces the same ``client_output`` records as ``Client.process_round`` (fused payloads) and falls back to that
generic path for anything it does not cover (non-SGD client optimizers, FedProx/FedLabels, local DP, quantization,
privacy metrics, personalization, text batches).
"""
from __future__ import annotations

import copy
import logging
import math
import time
from typing import Dict, List

import numpy as np
import torch

from ..ops import _ext, arena_ops
from ..parallel.arena import ArenaLayout, adopt_module, module_arena
from ..utils import print_rank
from ..utils.timing import PHASES
from . import client as client_mod

REC_LOSS, REC_SUM, REC_SUMSQ, REC_COUNT, REC_NS, REC_WEIGHT, REC_STEPS, REC_PAD = range(8)


class _LossModule(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, batch):
        return self.m.loss(batch)


class _Slot:
    def __init__(self, idx, model, device):
        self.idx = idx
        self.model = model
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self.graphs: Dict[tuple, object] = {}
        self.static: Dict[tuple, dict] = {}
        self.pool = None
        self.params = [p for p in model.parameters()]
        self.grad_views = module_arena(model)[1].views()
        self.kernels_per_replay: Dict[tuple, int] = {}


class DeferredRound:
    """Handle for a round whose GPU work is enqueued but whose per-client records have not been read back yet."""

    def __init__(self, n_clients, weight_sum, resolve):
        self.n_clients = n_clients
        self.weight_sum = weight_sum          # device scalar: Σ aggregation weights (what the fused update divides by)
        self._resolve = resolve
        self._outs = None

    def resolve(self):
        if self._outs is None:
            self._outs = self._resolve()
            self._resolve = None
        return self._outs


class DeviceClientEngine:
    @staticmethod
    def maybe_create(worker, config, task):
        ds = client_mod.train_dataset
        if ds is None or not hasattr(ds, "device_tensors") or not hasattr(ds, "transform_batch"):
            return None
        eng = DeviceClientEngine(worker, config, task)
        return eng if eng.supports(config) else None

    def __init__(self, worker, config, task):
        self.worker, self.config, self.task = worker, config, task
        b200 = config["server_config"].get("b200", {}) or {}
        self.device = next(worker.model.parameters()).device
        self.use_graphs = bool(b200.get("cuda_graphs", True)) and self.device.type == "cuda"
        self.resident = bool(b200.get("device_resident_data", True))
        self.want_wave = bool(b200.get("wave_batched", True))
        ncpi = config["server_config"]["num_clients_per_iteration"]
        ncpi = max(int(x) for x in ncpi.split(",")) if isinstance(ncpi, str) else int(ncpi)
        import os as _os
        world = int(_os.environ.get("WORLD_SIZE", 1))
        workers = world if (world == 1 or b200.get("server_is_worker", True)) else world - 1
        share = math.ceil(ncpi / max(workers, 1)) + (1 if workers > 1 else 0)   # LPT may give one rank an extra client
        self.S = max(1, min(int(b200.get("max_concurrent_clients", 16)), share))
        self.dataset = client_mod.train_dataset
        self._built = False
        self.h2d_bytes_last_round = 0
        self.d2h_bytes_last_round = 0
        self.wave_ok = None          # None = untried, True/False after the first attempt

    # ------------------------------------------------------------------ capability
    def supports(self, cfg) -> bool:
        cc, sc = cfg["client_config"], cfg["server_config"]
        strategy = str(cfg["strategy"]).lower()
        if strategy not in ("fedavg", "dga", "fedprox"):
            return False
        if cc["optimizer_config"].get("type") != "sgd":
            return False
        dp = cfg.get("dp_config", None) or {}
        # local DP (DGA): clip / normalise + Gaussian noise are fused into the gather; FedProx: closed-form proximal
        # gradient in the fused step; gradient quantization (DGA): radix-select statistics + binning fused into the gather.
        if cc.get("quant_thresh", None) is not None:
            if strategy != "dga":
                return False                   # only DGA quantizes in the reference (dga.py:149)
            if dp.get("enable_local_dp", False) and float(dp.get("eps", -1)) >= 0:
                return False                   # quantizing a noised gradient needs it materialised: generic path
        if dp.get("enable_local_dp", False) and strategy != "dga":
            return False
        pm = cfg.get("privacy_metrics_config", None) or {}
        if pm.get("apply_metrics", False) or sc.get("type", "model_optimization") == "personalization":
            return False
        if cc.get("annealing_config", None) is not None or cc.get("ss_config", None) is not None:
            return False
        if "updatable_names" in cc.get("trainer_config", {}) or cc.get("stats_on_smooth_grad", False):
            return False
        if cfg["model_config"].get("freeze_layer", None) or sc.get("send_dicts", False):
            return False
        if cc.get("ignore_subtask", False) and hasattr(self.worker.model, "single_task_loss"):
            return False
        return True

    # ------------------------------------------------------------------ lazy build
    def _build(self):
        base = self.worker.model
        ar = module_arena(base) or adopt_module(base, with_grad=True)
        self.layout: ArenaLayout = ar[0].layout
        P = self.layout.padded_numel
        dev = self.device
        opt = self.config["client_config"]["optimizer_config"]
        self.weight_decay = float(opt.get("weight_decay", 0.0) or 0.0)
        self._pack_dataset()
        self.slot_plan = self._plan_compact_slots(base)
        if self.slot_plan is not None:
            # compact slot arenas: only live parameters exist per client (see SlotBatchedResNet.plan_compact); the
            # per-slot nn.Modules of the generic fallback cannot alias such rows, so this mode is lock-step only
            P = int(self.slot_plan["numel"])
            self.index_map = self.slot_plan["index_map"].to(dev)
            print_rank("device engine: compact slot arenas, {} of {} parameters live per client ({} filters pruned)".format(
                int((self.index_map >= 0).sum()), self.layout.numel, len(self.slot_plan["compact"])), logging.INFO)
        else:
            self.index_map = None
        self.W = torch.zeros(self.S, P, device=dev)
        self.G = torch.zeros(self.S, P, device=dev)
        # slot-layout rows: the round's global weights and the round's accumulator (the permutation to / from the
        # global arena layout is paid once per round on these single rows, see csrc/gather_kernels.cu)
        self.wg_slot = torch.zeros(P, device=dev) if self.index_map is not None else None
        self.acc_slot = torch.zeros(P, device=dev) if self.index_map is not None else None
        self.pg_norm2 = torch.zeros(self.S, device=dev)
        self._dead_idx = None
        self._w_ref = None
        strategy = str(self.config["strategy"]).lower()
        cc = self.config["client_config"]
        self.prox_mult = self.prox_loss = None
        if strategy == "fedprox":
            mu = float(cc.get("mu", 0.001))
            ref_mult = bool(cc.get("fedprox_reference_multiplicity", True))
            lay = self.layout
            n_t = len(lay.sizes)
            glob = torch.zeros(lay.padded_numel)
            for i, (o, k) in enumerate(zip(lay.offsets, lay.sizes)):
                glob[o:o + k] = mu * ((n_t - i) if ref_mult else 1.0)     # the reference counts tensor i (P - i) times
            if self.index_map is not None:
                m = self.index_map.long().cpu()
                row = torch.where(m >= 0, glob[m.clamp(min=0)], torch.zeros(()))
            else:
                row = glob
            self.prox_mult = row.to(dev).contiguous()
            self.prox_loss = torch.zeros(self.S, device=dev)
        self.dp = dict(self.config.get("dp_config", None) or {})
        self.local_dp = bool(self.dp.get("enable_local_dp", False)) and strategy == "dga"
        self.slots: List[_Slot] = []
        for s in range(self.S if self.slot_plan is None else 0):
            m = copy.deepcopy(base)
            for attr in ("_flute_arena", "_flute_client_ctx"):
                if hasattr(m, attr):
                    delattr(m, attr)
            adopt_module(m, with_grad=True, param_buffer=self.W[s], grad_buffer=self.G[s])
            m.train()
            self.slots.append(_Slot(s, m, dev))
        self.hyper = arena_ops.make_hyper(self.S, dev)
        self.stats = torch.zeros(self.S, 4, device=dev)
        self.loss_sum = torch.zeros(self.S, device=dev)
        self.weights = torch.zeros(self.S, device=dev)
        self.active = torch.zeros(self.S, dtype=torch.int32, device=dev)
        opt = self.config["client_config"]["optimizer_config"]
        self.momentum = float(opt.get("momentum", 0.0) or 0.0)
        self.M = torch.zeros(self.S, P, device=dev) if self.momentum != 0 else None
        self.first = torch.ones(self.S, dtype=torch.int32, device=dev) if self.momentum != 0 else None
        self.nesterov = bool(opt.get("nesterov", False))
        self.dampening = float(opt.get("dampening", 0.0) or 0.0)
        # stacked (strided) per-parameter views over the slot arenas for the vmapped step
        names = [n for n, _ in base.named_parameters()]
        lay = self.layout
        self.has_buffers = any(True for _ in base.buffers())
        self.slot_model = None
        if self.slot_plan is not None and self.slot_plan.get("kind") == "slotnet":
            from ..models.slotnet_resnet import SlotNetResNet
            bs = int(self.config["client_config"]["data_config"]["train"]["batch_size"])
            self.slot_model = SlotNetResNet(base, self.W, self.G, self.slot_plan, batch=bs)
            self.param_stack, self.grad_stack, self.param_names, self.loss_module = {}, [], [], None
            print_rank("device engine: SlotNet (TMA + tcgen05 NHWC program, {} launches per local step) enabled".format(
                self.slot_model.n_ops), logging.INFO)
        elif self.slot_plan is not None:
            from ..models.slot_resnet import SlotBatchedResNet
            self.slot_model = SlotBatchedResNet(base, self.layout, self.W, self.G, plan=self.slot_plan)
            self.param_stack, self.grad_stack, self.param_names, self.loss_module = {}, [], [], None
            print_rank("device engine: slot-batched hand-written conv/GroupNorm path enabled", logging.INFO)
        else:
            self.param_stack = {"m." + n: self.W[:, o:o + k].view((self.S,) + tuple(sh))
                                for n, o, k, sh in zip(names, lay.offsets, lay.sizes, lay.shapes)}
            self.grad_stack = [self.G[:, o:o + k].view((self.S,) + tuple(sh))
                               for o, k, sh in zip(lay.offsets, lay.sizes, lay.shapes)]
            self.param_names = ["m." + n for n in names]
            self.loss_module = _LossModule(self.slots[0].model)
            if self.device.type == "cuda" and self.want_wave:
                from ..models.slot_resnet import SlotBatchedResNet
                ext = _ext.load()
                if ext is not None and hasattr(ext, "slot_conv_fprop") and SlotBatchedResNet.supports(base):
                    self.slot_model = SlotBatchedResNet(self.slots[0].model, self.layout, self.W, self.G)
                    print_rank("device engine: slot-batched hand-written conv/GroupNorm path enabled", logging.INFO)
        self.wave_graphs: Dict[tuple, object] = {}
        self.wave_static: Dict[tuple, dict] = {}
        self._staging: Dict[tuple, list] = {}
        self.wave_kernels: Dict[tuple, int] = {}
        self.wave_pool = None
        self._built = True

    def _plan_compact_slots(self, base):
        """Compact slot arenas are used when (a) the model has a slot-batched executor, (b) no weight decay touches
        parameters whose gradient is structurally zero, (c) every client runs the same number of full batches (all slots
        stay in lock-step, the per-slot fallback is never needed)."""
        import os as _os
        if self.device.type != "cuda" or not self.want_wave or _os.environ.get("FLUTE_COMPACT_SLOTS", "1") == "0":
            return None
        if self.weight_decay != 0.0:
            return None
        from ..models.slot_resnet import SlotBatchedResNet
        ext = _ext.load()
        if ext is None or not hasattr(ext, "slot_scatter_in") or not SlotBatchedResNet.supports(base):
            return None
        dcfg = self.config["client_config"]["data_config"]["train"]
        bs = int(dcfg["batch_size"])
        sizes = set(int(self.offsets[i + 1] - self.offsets[i]) for i in range(len(self.offsets) - 1))
        if len(sizes) != 1 or next(iter(sizes)) % bs != 0:
            return None
        src = self.X if self.resident else self.Xh
        example = self.dataset.transform_batch(src[:2].to(self.device))
        if _os.environ.get("FLUTE_SLOTNET", "1") != "0":
            from ..models.slotnet_resnet import SlotNetResNet
            if SlotNetResNet.supports(base, example):
                return SlotNetResNet.plan(base, self.layout)
        return SlotBatchedResNet.plan_compact(base, self.layout, example)

    def _pack_dataset(self):
        """One contiguous (pinned) pack of every user's raw samples + per-user offsets; uploaded once if resident."""
        t0 = time.time()
        ds = self.dataset
        xs, ys, off = [], [], [0]
        for u in ds.user_list:
            t = ds.device_tensors(u)
            xs.append(t["x"])
            ys.append(t["y"])
            off.append(off[-1] + int(t["x"].shape[0]))
        self.offsets = off
        X, Y = torch.cat(xs), torch.cat(ys)
        cuda = self.device.type == "cuda"
        if self.resident:
            self.X, self.Y = X.to(self.device), Y.to(self.device)
            self.Xh = self.Yh = None
        else:
            self.Xh, self.Yh = (X.pin_memory(), Y.pin_memory()) if cuda else (X, Y)
            max_n = max(b - a for a, b in zip(off[:-1], off[1:]))
            self.X = torch.empty((self.S * max_n,) + tuple(X.shape[1:]), dtype=X.dtype, device=self.device)
            self.Y = torch.empty((self.S * max_n,) + tuple(Y.shape[1:]), dtype=Y.dtype, device=self.device)
            self.stage_rows = max_n
        print_rank("device engine: packed {} users / {} samples ({:.1f} MB, {}) in {:.1f}s".format(
            len(ds.user_list), off[-1], X.numel() * X.element_size() / 2 ** 20,
            "HBM-resident" if self.resident else "pinned host, streamed per round", time.time() - t0), logging.INFO)

    def set_resident(self, resident: bool):
        """Switch between HBM-resident and streamed inputs (bench.py's end-to-end pass)."""
        if self._built and resident != self.resident:
            self.resident = resident
            self._pack_dataset()
        else:
            self.resident = resident

    def _slot_rows(self, s, cid):
        """(base_row, n) of client ``cid``'s samples for slot ``s`` — streaming copies the shard H2D first."""
        a, b = self.offsets[cid], self.offsets[cid + 1]
        n = b - a
        if self.resident:
            return a, n
        base = s * self.stage_rows
        self.X[base:base + n].copy_(self.Xh[a:b], non_blocking=True)
        self.Y[base:base + n].copy_(self.Yh[a:b], non_blocking=True)
        self.h2d_bytes_last_round += (b - a) * (self.Xh[0].numel() * self.Xh.element_size()
                                                + self.Yh[0].numel() * self.Yh.element_size())
        return base, n

    # ------------------------------------------------------------------ per-slot mini-batch
    def _step_body(self, slot: _Slot, xraw, y):
        batch = {"x": self.dataset.transform_batch(xraw), "y": y}
        loss = slot.model.loss(batch)
        grads = torch.autograd.grad(loss, slot.params, allow_unused=True)
        dst = [v for v, g in zip(slot.grad_views, grads) if g is not None]
        src = [g for g in grads if g is not None]
        torch._foreach_copy_(dst, src)
        s = slot.idx
        arena_ops.fused_client_step(
            self.W[s:s + 1], self.G[s:s + 1], self.hyper[s:s + 1], self.stats[s:s + 1],
            self.M[s:s + 1] if self.M is not None else None, n_logical=self.layout.numel, nesterov=self.nesterov,
            dampening=self.dampening, zero_grad=True, first_step=self.first[s:s + 1] if self.first is not None else None,
            prox_ref=self._prox_ref(), prox_mult=self.prox_mult,
            prox_loss=self.prox_loss[s:s + 1] if self.prox_loss is not None else None)
        if self.first is not None:
            self.first[s:s + 1].zero_()
        self.loss_sum[s:s + 1] += loss.detach().float().reshape(1)
        if self.prox_loss is not None:
            self.loss_sum[s:s + 1] += self.prox_loss[s:s + 1]
            self.prox_loss[s:s + 1].zero_()

    def _run_step(self, slot: _Slot, idx):
        key = int(idx.numel())
        if not self.use_graphs:
            self._step_body(slot, self.X.index_select(0, idx), self.Y.index_select(0, idx))
            return
        st = slot.static.get(key)
        if st is None:
            st = {"x": torch.empty((key,) + tuple(self.X.shape[1:]), dtype=self.X.dtype, device=self.device),
                  "y": torch.empty((key,) + tuple(self.Y.shape[1:]), dtype=self.Y.dtype, device=self.device), "uses": 0}
            slot.static[key] = st
        torch.index_select(self.X, 0, idx, out=st["x"])
        torch.index_select(self.Y, 0, idx, out=st["y"])
        g = slot.graphs.get(key)
        if g is None:
            st["uses"] += 1
            if st["uses"] <= 2:          # warm up eagerly (cuDNN/cuBLAS handles + autotune on this stream)
                self._step_body(slot, st["x"], st["y"])
                return
            g = torch.cuda.CUDAGraph()
            n0 = _ext.LAUNCH_COUNTER["n"]
            # thread_local: the async checkpoint writer thread may issue D2H copies while we capture
            with torch.cuda.graph(g, stream=slot.stream, pool=slot.pool,   # one private pool PER SLOT:
                                  capture_error_mode="thread_local"):
                self._step_body(slot, st["x"], st["y"])                     # slots replay concurrently
            if slot.pool is None:
                slot.pool = g.pool()
            slot.kernels_per_replay[key] = _ext.LAUNCH_COUNTER["n"] - n0
            slot.graphs[key] = g
        g.replay()
        _ext.count_launch(slot.kernels_per_replay.get(key, 0))

    # ------------------------------------------------------------------ wave-batched mini-batch (all slots at once)
    def _wave_body(self, xw, yw):
        if self.slot_model is not None:
            S, B = xw.shape[0], xw.shape[1]
            xb = self.dataset.transform_batch(xw.reshape((S * B,) + tuple(xw.shape[2:])))
            if hasattr(self.slot_model, "step"):       # SlotNet: static forward+backward program, no autograd
                losses = self.slot_model.step(xb, yw)
            else:
                losses = self.slot_model.losses(xb.reshape((S, B) + tuple(xb.shape[1:])), yw)
                losses.sum().backward()                # weight grads are accumulated into self.G by the kernels
            arena_ops.fused_client_step(self.W, self.G, self.hyper, self.stats, self.M, n_logical=self.layout.numel,
                                        nesterov=self.nesterov, dampening=self.dampening, zero_grad=True,
                                        first_step=self.first, prox_ref=self._prox_ref(), prox_mult=self.prox_mult,
                                        prox_loss=self.prox_loss)
            if self.first is not None:
                self.first.zero_()
            self.loss_sum += losses.detach().float()
            if self.prox_loss is not None:             # the reference's batch loss includes the proximal term
                self.loss_sum += self.prox_loss
                self.prox_loss.zero_()
            return
        from torch.func import functional_call, grad_and_value, vmap

        def loss_fn(p, xb, yb):
            return functional_call(self.loss_module, p, ({"x": self.dataset.transform_batch(xb), "y": yb},))

        grads, losses = vmap(grad_and_value(loss_fn), in_dims=(0, 0, 0), randomness="different")(self.param_stack, xw, yw)
        torch._foreach_copy_(self.grad_stack, [grads[n] for n in self.param_names])
        arena_ops.fused_client_step(self.W, self.G, self.hyper, self.stats, self.M, n_logical=self.layout.numel,
                                    nesterov=self.nesterov, dampening=self.dampening, zero_grad=True,
                                    first_step=self.first, prox_ref=self._prox_ref(), prox_mult=self.prox_mult,
                                    prox_loss=self.prox_loss)
        if self.first is not None:
            self.first.zero_()
        self.loss_sum += losses.detach().float()
        if self.prox_loss is not None:
            self.loss_sum += self.prox_loss
            self.prox_loss.zero_()

    def _prox_ref(self):
        """FedProx anchor = the round's global weights in the slot layout (a static buffer, safe inside CUDA graphs)."""
        if self.prox_mult is None:
            return None
        return self.wg_slot if self.wg_slot is not None else self._w_ref

    def _quant_tables(self):
        """(segs [T, 3], seg_of_blk [P / 32]) of the slot layout for the fused quantization kernels."""
        if getattr(self, "_qtab", None) is None:
            lay = self.layout
            P = int(self.W.shape[1])
            names = lay.names
            if self.slot_plan is not None:
                offs = [int(self.slot_plan["offsets"][n]) for n in names]
                order = sorted(range(len(offs)), key=lambda i: offs[i])
                ends = {order[k]: (offs[order[k + 1]] if k + 1 < len(order) else P) for k in range(len(order))}
                m = self.index_map.cpu()
                # iterated span = up to the tensor's last stored element (internal layout padding is iterated as zeros;
                # the kernels correct the zero count with total - span, which may be negative)
                live = []
                for i in range(len(offs)):
                    pos = (m[offs[i]:ends[i]] >= 0).nonzero()
                    live.append(int(pos.max()) + 1 if pos.numel() else 0)
            else:
                offs, live = list(lay.offsets), list(lay.sizes)
            segs = torch.tensor([[o, n, int(nt)] for o, n, nt in zip(offs, live, lay.sizes)], dtype=torch.int64)
            blk = torch.full((P // 32,), -1, dtype=torch.int16)
            for t, (o, n) in enumerate(zip(offs, live)):
                blk[o // 32:(o + n + 31) // 32] = t
            self._qtab = (segs.to(self.device), blk.to(self.device))
        return self._qtab

    def _dead_coords(self):
        """Global-arena positions of real parameters that no slot stores (elided dead filter taps)."""
        if self._dead_idx is None:
            lay = self.layout
            is_param = torch.zeros(lay.padded_numel, dtype=torch.bool)
            for o, k in zip(lay.offsets, lay.sizes):
                is_param[o:o + k] = True
            m = self.index_map.long().cpu()
            is_param[m[m >= 0]] = False
            self._dead_idx = is_param.nonzero().view(-1).to(torch.int32).to(self.device)
        return self._dead_idx

    def _run_wave_step(self, idx2d):
        """``idx2d``: [S, B] global sample rows.  Returns False if this model cannot be vmapped (caller falls back)."""
        B = int(idx2d.shape[1])
        st = self.wave_static.get(B)
        if st is None:
            st = {"x": torch.empty((self.S, B) + tuple(self.X.shape[1:]), dtype=self.X.dtype, device=self.device),
                  "y": torch.empty((self.S, B) + tuple(self.Y.shape[1:]), dtype=self.Y.dtype, device=self.device), "uses": 0}
            self.wave_static[B] = st
        flat = idx2d.reshape(-1)
        torch.index_select(self.X, 0, flat, out=st["x"].view((self.S * B,) + tuple(self.X.shape[1:])))
        torch.index_select(self.Y, 0, flat, out=st["y"].view((self.S * B,) + tuple(self.Y.shape[1:])))
        g = self.wave_graphs.get(B)
        if g is None:
            st["uses"] += 1
            if st["uses"] <= 2 or not self.use_graphs:
                try:
                    self._wave_body(st["x"], st["y"])
                except Exception as e:  # no vmap rule for some op in this model → per-slot path from now on
                    if self.wave_ok is None:
                        print_rank("wave-batched step unavailable for this model ({}: {}); using per-slot graphs"
                                   .format(type(e).__name__, str(e).split("\n")[0][:160]), logging.WARNING)
                        self.wave_ok = False
                        return False
                    raise
                self.wave_ok = True
                return True
            g = torch.cuda.CUDAGraph()
            n0 = _ext.LAUNCH_COUNTER["n"]
            with torch.cuda.graph(g, pool=self.wave_pool, capture_error_mode="thread_local"):
                self._wave_body(st["x"], st["y"])
            if self.wave_pool is None:
                self.wave_pool = g.pool()
            self.wave_kernels[B] = _ext.LAUNCH_COUNTER["n"] - n0
            self.wave_graphs[B] = g
        g.replay()
        _ext.count_launch(self.wave_kernels.get(B, 0))
        return True

    # ------------------------------------------------------------------ a round's share
    def train_clients(self, client_ids, lr, nround, w_global, acc, defer=False):
        """Train ``client_ids`` (all enqueued, no host sync inside).  ``defer=True`` returns a :class:`DeferredRound`:
        the per-client record table is copied to pinned memory asynchronously and only read when the caller
        ``resolve()``s it — the server enqueues its fused update / checkpoint snapshot first, so the host work of a
        round overlaps the GPU instead of following it."""
        if not self._built:
            self._build()
        cfg = self.config
        dcfg = cfg["client_config"]["data_config"]["train"]
        bs = int(dcfg["batch_size"])
        desired = dcfg.get("desired_max_samples", None)
        max_norm = dcfg.get("max_grad_norm", None)
        strategy = str(cfg["strategy"]).lower()
        sc = cfg["server_config"]
        softmax = strategy == "dga" and sc.get("aggregate_median", None) == "softmax"
        dev = self.device
        self.h2d_bytes_last_round = 0
        if not torch.is_tensor(w_global):
            w_global = self.worker.weight_buffer()     # Worker.set_weights already packed the list into it
        t_begin = time.time()
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        records = torch.zeros(len(client_ids), 8, device=dev)
        hyper_row = torch.tensor([lr, max_norm or 0.0, self.weight_decay, self.momentum], dtype=torch.float32)
        self.hyper.copy_(hyper_row.to(dev, non_blocking=True).expand(self.S, 4))
        wave_mode = self.want_wave and (self.slot_model is not None or not self.has_buffers) and self.wave_ok is not False
        for wave_start in range(0, len(client_ids), self.S):
            wave = client_ids[wave_start:wave_start + self.S]
            n_act = len(wave)
            with PHASES.phase("bcast_local"):
                if self.index_map is not None:
                    # the "broadcast" into every (compact) slot: one gather through the layout permutation, S + 1 rows out
                    arena_ops.slot_gather_bcast(self.W, self.wg_slot, w_global, self.index_map)
                else:
                    self.W.copy_(w_global.view(1, -1).expand(self.S, -1))   # the "broadcast" into every slot
                    if self.prox_mult is not None:
                        if getattr(self, "_w_ref", None) is None:
                            self._w_ref = torch.empty_like(w_global)
                        self._w_ref.copy_(w_global)
            self.stats.zero_()
            self.loss_sum.zero_()
            if self.first is not None:
                self.first.fill_(1)
            rows = [self._slot_rows(s, cid) for s, cid in enumerate(wave)]
            ns = [n for _, n in rows]
            nbs = []
            for n in ns:
                nb = math.ceil(n / bs)
                if desired is not None:
                    nb = min(nb, max(1, math.ceil(desired / bs)))
                nbs.append(nb)
            # on-device shuffles: one [S, n] argsort when all users have the same size, else per-slot randperm
            if len(set(ns)) == 1:
                bases = [b for b, _ in rows] + [rows[0][0]] * (self.S - n_act)   # idle slots re-read slot 0's shard
                perm = torch.rand(self.S, ns[0], device=dev).argsort(dim=1)
                perm = perm + torch.tensor(bases, device=dev).view(-1, 1)
                perms = None
            else:
                perm = None
                perms = [torch.randperm(n, device=dev) + b for b, n in rows]
            full_steps = min(min(nb, n // bs) for n, nb in zip(ns, nbs))   # steps every slot runs with a full batch
            ns_done = [0] * n_act
            b0 = 0
            if wave_mode and perm is not None and full_steps > 0:
                for b in range(full_steps):
                    if not self._run_wave_step(perm[:, b * bs:(b + 1) * bs]):
                        wave_mode = False
                        break
                    b0 = b + 1
                    for s in range(n_act):
                        ns_done[s] += bs
            if b0 < max(nbs) and self.slot_plan is not None:
                raise RuntimeError("compact slot arenas require lock-step clients (equal shard sizes, full batches); "
                                   "set FLUTE_COMPACT_SLOTS=0")
            if b0 < max(nbs):
                # per-slot path for whatever could not run in lock-step
                if main is not None:
                    for slot in self.slots[:n_act]:
                        slot.stream.wait_stream(main)
                for b in range(b0, max(nbs)):
                    for s in range(n_act):
                        if b >= nbs[s]:
                            continue
                        p = perm[s] if perm is not None else perms[s]
                        idx = p[b * bs:min((b + 1) * bs, ns[s])]
                        slot = self.slots[s]
                        if slot.stream is not None:
                            with torch.cuda.stream(slot.stream):
                                self._run_step(slot, idx)
                        else:
                            self._run_step(slot, idx)
                        ns_done[s] += int(idx.numel())
                if main is not None:
                    for slot in self.slots[:n_act]:
                        main.wait_stream(slot.stream)
            # aggregation weights (device side, no sync)
            ns_t = torch.tensor(ns_done + [0] * (self.S - n_act), dtype=torch.float32).to(dev, non_blocking=True)
            if softmax:
                kind = sc.get("weight_train_loss", "train_loss")
                mean_, mag_, var_, _ = arena_ops.finalize_stats(self.stats)
                signal = {"train_loss": self.loss_sum / ns_t.clamp(min=1), "mag_var_loss": var_,
                          "mag_mean_loss": mean_}.get(kind, mag_)
                w = torch.exp(-float(sc["softmax_beta"]) * signal)
                w = torch.where(torch.isfinite(w), w, torch.zeros_like(w)).clamp(max=100.0)
            elif strategy == "dga":
                w = torch.ones(self.S, device=dev)
            else:
                w = ns_t.clone()
            act = torch.zeros(self.S, dtype=torch.int32)
            act[:n_act] = 1
            self.active.copy_(act.to(dev, non_blocking=True))
            actf = self.active.float()
            wg_row = self.wg_slot if self.index_map is not None else w_global
            acc_row = self.acc_slot if self.index_map is not None else acc
            sig = seeds = None
            with PHASES.phase("gather_local"):
                if self.local_dp:
                    # local DP fused into the gather (ref. extensions/privacy/__init__.py:154-201): per-client norm of the
                    # pseudo-gradient -> clip (eps < 0) or normalise to max_grad + Gaussian noise; all on the device
                    dpc = self.dp
                    C = float(dpc["max_grad"])
                    norm = arena_ops.slot_pg_sqnorm(self.W, wg_row, self.pg_norm2).sqrt().clamp(min=1e-12)
                    if float(dpc["eps"]) < 0:
                        scale = (C / norm).clamp(max=1.0)
                    else:
                        from ..extensions.privacy import compute_LDP_noise_std, rng as dp_rng
                        scaler = float(dpc.get("weight_scaler", 1))
                        sens = math.sqrt(C ** 2 + (float(dpc["max_weight"]) ** 2 if softmax else 0.0))
                        sigma = float(compute_LDP_noise_std(float(dpc["eps"]), sens, float(dpc.get("delta", 1e-7))))
                        scale = C / norm
                        if softmax:                                    # noisy aggregation weight (add_weight_noise)
                            noisy = (w * scaler).clamp(max=float(dpc["max_weight"])) + sigma * dp_rng.randn((self.S,), dev)
                            w = noisy.clamp(min=float(dpc["min_weight"]), max=float(dpc["max_weight"])) / scaler
                        sig = w * actf * sigma
                        seeds = torch.tensor([dp_rng.dp_seed(stream=100 + i) for i in range(self.S)],
                                             dtype=torch.int64).to(dev, non_blocking=True)
                    self.weights.copy_(w * actf)
                    coef = self.weights * scale
                else:
                    self.weights.copy_(w * actf)
                    coef = self.weights
                q_thresh = cfg["client_config"].get("quant_thresh", None)
                if q_thresh is not None and strategy == "dga":
                    # gradient quantization of every client's payload (ref. quant.py), statistics by radix select
                    bits = int(cfg["client_config"].get("quant_bits", 8))
                    segs, seg_of_blk = self._quant_tables()
                    qparams = arena_ops.slot_quant_stats(self.W, wg_row, segs, float(q_thresh), bits)
                    arena_ops.slot_quant_gather(acc_row, self.W, wg_row, coef, qparams, seg_of_blk, bits)
                else:
                    arena_ops.slot_gather_fused(acc_row, self.W, wg_row, coef, sig, seeds)
                if self.index_map is not None:
                    arena_ops.slot_scatter_acc(acc, self.acc_slot, self.index_map)
                    if sig is not None:
                        from ..extensions.privacy import rng as dp_rng
                        arena_ops.dead_coord_noise(acc, self._dead_coords(), (sig * sig).sum(), dp_rng.dp_seed(stream=99))
            rec = records[wave_start:wave_start + n_act]
            rec[:, REC_LOSS] = self.loss_sum[:n_act]
            rec[:, REC_SUM:REC_COUNT + 1] = self.stats[:n_act, 0:3]
            rec[:, REC_NS] = ns_t[:n_act]
            rec[:, REC_WEIGHT] = self.weights[:n_act]
        self.d2h_bytes_last_round = records.numel() * records.element_size()
        n_clients = len(client_ids)

        def build(host):
            t_end = time.time()
            per = (t_end - t_begin) / max(n_clients, 1)
            outs = []
            for i in range(n_clients):
                r = host[i]
                n = max(r[REC_COUNT], 1.0)
                mag = math.sqrt(max(r[REC_SUMSQ], 0.0) / n)
                outs.append({
                    "cs": {"setup": 0.0, "training": per, "full cost": per},
                    "tl": float(r[REC_LOSS]), "mg": np.float32(mag), "vg": np.float32(r[REC_SUMSQ] / n - mag * mag),
                    "ng": np.float32(r[REC_SUM] / n), "rg": np.float32(math.sqrt(max(r[REC_SUMSQ], 0.0))),
                    "ns": int(r[REC_NS]), "pl": {"weight": float(r[REC_WEIGHT]), "gradients": None, "fused": True},
                    "ts": t_end,
                })
            return outs

        if defer and dev.type == "cuda":
            pinned = self._record_staging(records.shape)
            pinned.copy_(records, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            wsum = records[:, REC_WEIGHT].sum()

            def resolve():
                ev.synchronize()                                         # the round's single device→host read
                return build(pinned.numpy().astype(np.float64))

            return DeferredRound(n_clients, wsum, resolve)
        return build(records.cpu().numpy().astype(np.float64))          # the round's single device→host read

    def _record_staging(self, shape):
        """Two alternating pinned buffers (a deferred round may still be unread when the next one is enqueued)."""
        key = tuple(shape)
        ring = self._staging.setdefault(key, [])
        if len(ring) < 2:
            ring.append(torch.empty(key, dtype=torch.float32).pin_memory())
            return ring[-1]
        ring.append(ring.pop(0))
        return ring[-1]
