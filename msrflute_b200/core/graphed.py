"""CUDA-graphed client mini-batch step for the generic (non-engine) client path.

The reference runs every mini-batch as ~10^2..10^3 eager launches plus a host sync (`/root/reference/core/trainer.py:
380-436`: ``loss.item()``, per-parameter gradient copies).  Configurations the slot engine declines (strategies that need
individual payloads, exotic models) still go through :class:`core.trainer.Trainer`; this module captures that trainer's
whole step — forward, backward, fused clip + statistics + SGD + zero-grad over the arena, loss accumulation — ONCE per
batch signature into a CUDA graph and replays it for every following mini-batch of every following client:

* inputs live in static device buffers (one ``copy_`` per tensor per step, the only eager work);
* parameters / gradients are the model's flat arenas, whose addresses never change between clients (receiving the
  global model is a copy INTO the arena), so one capture serves the whole job;
* hyper-parameters (lr, clip, weight decay, momentum) are read by the kernels from a device tensor, so learning-rate
  schedules do not invalidate the capture.

Anything that cannot be captured (a model with host-side control flow, a stock ``torch.optim`` optimizer, CPU tensors)
makes :meth:`GraphedTrainStep.__call__` fall back to the eager step — permanently for that signature — so enabling the
feature is always safe.  Enable with ``client_config.graphed_step: true`` or ``FLUTE_GRAPHED_STEP=1``.
"""
from __future__ import annotations

import logging
import os

import torch

logger = logging.getLogger(__name__)


def enabled(client_config=None) -> bool:
    env = os.environ.get("FLUTE_GRAPHED_STEP")
    if env is not None:
        return env == "1"
    if client_config is None:
        return False
    get = client_config.get if hasattr(client_config, "get") else lambda k, d=None: getattr(client_config, k, d)
    return bool(get("graphed_step", False))


def _signature(batch):
    sig = []
    for k in sorted(batch.keys()):
        v = batch[k]
        if torch.is_tensor(v):
            sig.append((k, tuple(v.shape), str(v.dtype)))
        else:
            sig.append((k, type(v).__name__, repr(v) if isinstance(v, (int, float, str, bool, type(None))) else id(type(v))))
    return tuple(sig)


class _Captured:
    __slots__ = ("graph", "static", "loss")

    def __init__(self, graph, static, loss):
        self.graph, self.static, self.loss = graph, static, loss


class GraphedTrainStep:
    """Callable replacement for ``Trainer._train_step`` (``trainer.step_fn``)."""

    WARMUP = 2            # eager steps on the side stream before capturing (cuBLAS workspaces, autotuners, lazy init)
    MAX_SIGNATURES = 4    # full batch + tail batch (+ slack); more distinct shapes -> eager

    def __init__(self, trainer):
        self.trainer = trainer
        self.cache = {}
        self.disabled = set()
        self.replays = 0
        self.captures = 0
        self.fallbacks = 0

    # ------------------------------------------------------------------------------------------------------------
    def capturable(self) -> bool:
        t = self.trainer
        if not torch.cuda.is_available() or not t.use_arena:
            return False
        try:
            dev = next(t.model.parameters()).device
        except StopIteration:
            return False
        if dev.type != "cuda" or t._arena() is None:
            return False
        if t.ss_scheduler is not None:        # scheduled sampling draws host-side random numbers every step
            return False
        from .trainer import _fused_adamw, _plain_sgd
        return t.optimizer is None or _plain_sgd(t.optimizer) or _fused_adamw(t.optimizer)

    def _to_static(self, batch, dev):
        static = {}
        for k, v in batch.items():
            static[k] = v.to(dev, non_blocking=True).clone() if torch.is_tensor(v) else v
        return static

    def _capture(self, batch):
        t = self.trainer
        dev = next(t.model.parameters()).device
        static = self._to_static(batch, dev)
        w, g = t._arena()
        # the warm-up steps must not change the training trajectory: snapshot and restore everything they touch
        hyper, stats, loss_sum = t.device_state()
        keep = [x.clone() for x in (w.flat, g.flat, stats, loss_sum)]
        mom = getattr(t, "_mom_state", None)
        keep_mom = [x.clone() for x in mom] if mom is not None else None
        adam = getattr(t, "_adam_state", None)
        keep_adam = [x.clone() for x in adam] if adam is not None else None
        buffers = [b for b in t.model.buffers()]
        keep_buf = [b.clone() for b in buffers]
        rng = torch.cuda.get_rng_state(dev)

        def restore():
            for dst, src in zip((w.flat, g.flat, stats, loss_sum), keep):
                dst.copy_(src)
            if keep_mom is not None and getattr(t, "_mom_state", None) is not None:
                for dst, src in zip(t._mom_state, keep_mom):
                    dst.copy_(src)
            for dst, src in zip(buffers, keep_buf):
                dst.copy_(src)
            if getattr(t, "_adam_state", None) is not None:
                for i, dst in enumerate(t._adam_state):
                    dst.copy_(keep_adam[i]) if keep_adam is not None else dst.zero_()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            try:
                for _ in range(self.WARMUP):
                    t._train_step(static)
            finally:
                restore()
        torch.cuda.current_stream(dev).wait_stream(side)
        if keep_mom is None and getattr(t, "_mom_state", None) is not None:
            t._mom_state[0].zero_()
            t._mom_state[1].fill_(1)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = t._train_step(static)
        # capture does not execute: state is still the restored one; the RNG offset consumed by the warm-up is irrelevant
        torch.cuda.set_rng_state(rng, dev)
        self.captures += 1
        return _Captured(graph, static, loss)

    def __call__(self, batch):
        t = self.trainer
        grp = t.optimizer.param_groups[0] if t.optimizer is not None else {}
        sig = _signature(batch) + ((id(t.optimizer), bool(grp.get("momentum", 0)), bool(grp.get("nesterov", False)),
                                    t.model.training),)
        if sig in self.disabled or not self.capturable():
            self.fallbacks += 1
            return t._train_step(batch)
        cap = self.cache.get(sig)
        if cap is None:
            if len(self.cache) >= self.MAX_SIGNATURES:
                self.disabled.add(sig)
                self.fallbacks += 1
                return t._train_step(batch)
            try:
                cap = self._capture(batch)
            except Exception as exc:           # noqa: BLE001 - any capture failure means "run eagerly", never "crash"
                logger.warning("graphed client step: capture failed for %s (%s: %s); running eagerly",
                               [s_[0] for s_ in sig[:-1]], type(exc).__name__, exc)
                torch.cuda.synchronize()
                self.disabled.add(sig)
                self.fallbacks += 1
                return t._train_step(batch)
            self.cache[sig] = cap
        for k, v in batch.items():
            if torch.is_tensor(v):
                cap.static[k].copy_(v, non_blocking=True)
        cap.graph.replay()
        self.replays += 1
        return cap.loss


def attach(ctx, trainer, client_config=None):
    """Give ``trainer`` the context's graphed step (created on first use, rebound when the trainer object changes)."""
    if not enabled(client_config):
        trainer.step_fn = None
        return None
    gs = ctx.graphed
    if gs is None or gs.trainer is not trainer:
        gs = GraphedTrainStep(trainer)
        ctx.graphed = gs
    trainer.step_fn = gs
    return gs
