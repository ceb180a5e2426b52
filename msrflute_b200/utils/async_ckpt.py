"""Asynchronous, coalescing checkpoint writer.

The reference ``torch.save``s ``latest_model.tar`` synchronously every round (``core/server.py:541-545``) — for
ResNet-18 that is a 45 MB device→host copy plus a disk write on the critical path of every round.  Here the round
only pays for a device-side snapshot (ONE D2D copy per underlying storage — the whole parameter arena at once — on a
side stream); a background thread moves the snapshot to pinned host memory (one D2H per storage) and writes it.  If rounds complete faster than the disk, queued
snapshots for the same path are coalesced (latest wins) — the file on disk is always a complete, consistent
checkpoint of some recent round, and ``flush()`` (called at shutdown, before any checkpoint is read back and by
``resume``) guarantees the final state is durable.
"""
from __future__ import annotations

import os
import queue
import threading

import torch


def _walk(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _walk(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_walk(v, fn) for v in obj)
    return obj


def _storage_key(t):
    return (t.device, t.untyped_storage().data_ptr())


def _snapshot(obj, stream):
    """Device-side copy of every tensor in ``obj``.  Tensors that are views into one big storage (the flat parameter
    arena: every entry of ``model.state_dict()``) are snapshotted with ONE copy of that storage and rebuilt as views of
    the copy — a round pays one D2D launch instead of one per parameter tensor."""
    groups = {}

    def collect(t):
        if t.is_cuda:
            g = groups.setdefault(_storage_key(t), [t.untyped_storage(), 0])
            g[1] += t.numel() * t.element_size()
        return t

    _walk(obj, collect)
    copies = {}
    ctx = torch.cuda.stream(stream) if stream is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        for key, (storage, used) in groups.items():
            if used * 2 >= storage.nbytes():            # mostly covered: copy the storage once
                whole = torch.empty(0, dtype=torch.uint8, device=key[0]).set_(storage)
                copies[key] = whole.clone().untyped_storage()

        def snap(t):
            t = t.detach()
            if t.is_cuda and _storage_key(t) in copies:
                return torch.empty(0, dtype=t.dtype, device=t.device).set_(
                    copies[_storage_key(t)], t.storage_offset(), t.size(), t.stride())
            return t.clone()

        return _walk(obj, snap)
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)


_PINNED = {}


def _to_host(obj):
    """D2H of a snapshot: one pinned staging buffer per distinct device storage.  Buffers are pooled by
    ``(nbytes, ordinal)`` — a checkpoint usually holds several storages of the SAME size (the parameter arena and the
    Adam ``m`` / ``v`` arenas, or ``exp_avg`` / ``exp_avg_sq`` of every same-shaped layer), and each needs its own
    staging buffer for the lifetime of the ``torch.save`` call.  The single writer thread reuses the pool between calls."""
    hosts = {}
    used = {}                              # nbytes -> how many buffers of that size this call already took

    def move(t):
        if not t.is_cuda:
            return t
        key = _storage_key(t)
        if key not in hosts:
            storage = t.untyped_storage()
            n = storage.nbytes()
            ordinal = used.get(n, 0)
            used[n] = ordinal + 1
            buf = _PINNED.get((n, ordinal))
            if buf is None:
                try:
                    buf = torch.empty(n, dtype=torch.uint8).pin_memory()
                except RuntimeError:
                    buf = torch.empty(n, dtype=torch.uint8)
                _PINNED[(n, ordinal)] = buf
            buf.copy_(torch.empty(0, dtype=torch.uint8, device=t.device).set_(storage), non_blocking=False)
            hosts[key] = buf.untyped_storage()
        return torch.empty(0, dtype=t.dtype).set_(hosts[key], t.storage_offset(), t.size(), t.stride())

    return _walk(obj, move)


class AsyncCheckpointer:
    """``min_interval``: seconds between two writes of the same path.  Rounds of the flagship take ~20 ms while pickling
    + writing a 47 MB checkpoint takes longer; a writer that runs back to back competes with the training thread for
    the GIL (every torch call re-acquires it).  Snapshots are still taken every round (latest wins); the file on disk
    is at most ``min_interval`` (+ one write) stale, and ``flush()`` always writes the newest snapshot."""

    def __init__(self, min_interval: float = 0.5):
        self.min_interval = float(os.environ.get("FLUTE_CKPT_MIN_INTERVAL", min_interval))
        import sys
        sys.setswitchinterval(min(sys.getswitchinterval(), 0.0005))   # bound GIL hand-over latency to the trainer
        self._last_write = {}              # path -> time of the last completed write
        self._flushing = 0
        self._pending = {}                 # path -> (state | text, event)
        self._lock = threading.Lock()
        self._wake = threading.Condition(self._lock)
        self._busy = 0
        self._stop = False
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._thread = threading.Thread(target=self._loop, name="flute-ckpt", daemon=True)
        self._thread.start()
        self.written = 0
        self.coalesced = 0

    def submit(self, path: str, state):
        ev = None
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            snap = _snapshot(state, self._stream)
            ev = torch.cuda.Event()
            ev.record(self._stream)
            # the next in-place weight update (this stream, or a peer's P2P stores ordered behind it) must not start
            # before the D2D snapshot has been taken
            torch.cuda.current_stream().wait_event(ev)
        else:
            snap = _snapshot(state, None)
        with self._wake:
            if path in self._pending:
                self.coalesced += 1
            self._pending[path] = (snap, ev)
            self._wake.notify()

    def submit_copy(self, src: str, dst: str):
        """Copy ``src`` to ``dst`` off the training thread (the periodic ``epoch<i>_best_*`` backups: three 47 MB file
        copies every ``model_backup_freq`` rounds would otherwise stall that round by ~70 ms).  A pending write of
        ``src`` is flushed first, so the copy is of the newest snapshot."""
        with self._wake:
            self._pending[dst] = (("copy", src), None)
            self._wake.notify()

    def submit_text(self, path: str, text: str):
        """Small text files (status_log.json, config.yaml): same latest-wins queue, written off the training thread."""
        with self._wake:
            if path in self._pending:
                self.coalesced += 1
            self._pending[path] = (text, None)
            self._wake.notify()

    def _due(self, now):
        """First pending path whose rate limit has expired (all of them while flushing / stopping)."""
        soonest = None
        for path in self._pending:
            wait = self._last_write.get(path, -1e30) + self.min_interval - now
            if wait <= 0 or self._flushing or self._stop:
                return path, 0.0
            soonest = wait if soonest is None else min(soonest, wait)
        return None, soonest

    def _loop(self):
        import time
        while True:
            with self._wake:
                while True:
                    if self._stop and not self._pending:
                        return
                    path, wait = self._due(time.monotonic())
                    if path is not None:
                        break
                    self._wake.wait(timeout=wait)
                snap, ev = self._pending.pop(path)
                self._busy += 1
            try:
                tmp = path + ".tmp"
                if isinstance(snap, tuple) and snap and snap[0] == "copy":
                    src = snap[1]
                    with self._wake:
                        queued = self._pending.pop(src, None)
                    if queued is not None:                 # newest snapshot of the source first
                        q_snap, q_ev = queued
                        if isinstance(q_snap, str):
                            with open(src + ".tmp", "w", encoding="utf8") as f:
                                f.write(q_snap)
                        else:
                            if q_ev is not None:
                                q_ev.synchronize()
                            torch.save(_to_host(q_snap), src + ".tmp")
                        os.replace(src + ".tmp", src)
                    if os.path.exists(src):
                        import shutil
                        shutil.copyfile(src, tmp)
                    else:
                        continue
                elif isinstance(snap, str):
                    with open(tmp, "w", encoding="utf8") as f:
                        f.write(snap)
                else:
                    import time as _t
                    t0 = _t.perf_counter()
                    if ev is not None:
                        ev.synchronize()
                    t1 = _t.perf_counter()
                    host = _to_host(snap)
                    t2 = _t.perf_counter()
                    torch.save(host, tmp)
                    t3 = _t.perf_counter()
                    if os.environ.get("FLUTE_CKPT_TRACE") == "1":
                        print("[ckpt] {} wait {:.1f} ms, d2h {:.1f} ms, torch.save {:.1f} ms at t={:.3f}".format(
                            os.path.basename(path), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, t3), flush=True)
                os.replace(tmp, path)
                self._last_write[path] = time.monotonic()
                self.written += 1
            except Exception as e:  # never kill training because of a checkpoint
                print("async checkpoint to {} failed: {}".format(path, e))
            finally:
                with self._wake:
                    self._busy -= 1
                    self._wake.notify_all()

    def flush(self):
        with self._wake:
            self._flushing += 1
            self._wake.notify_all()
            try:
                while self._pending or self._busy:
                    self._wake.wait(timeout=0.05)
            finally:
                self._flushing -= 1

    def close(self):
        self.flush()
        with self._wake:
            self._stop = True
            self._wake.notify_all()


_CKPT = None


def get_checkpointer() -> AsyncCheckpointer:
    global _CKPT
    if _CKPT is None:
        _CKPT = AsyncCheckpointer()
    return _CKPT


def flush_checkpoints():
    if _CKPT is not None:
        _CKPT.flush()
