"""Device-side timing, NVTX ranges and optional Kineto traces (the reference has cProfile and wall-clock
round stats only).

* :class:`RoundTimer` — one CUDA event per round on the training stream; elapsed times are read lazily with
  ``event.query()`` so the host never blocks.  The server logs them as ``devMsPerRound`` next to the reference's
  wall-clock ``secsPerRoundTotal``.
* :func:`nvtx_range` — ``with nvtx_range("clients"):`` annotates phases for nsys/ncu when ``FLUTE_NVTX=1``.
* :class:`TraceWindow` — ``FLUTE_TRACE=/path/trace.json[:first[:count]]`` captures rounds ``[first, first+count)`` with
  ``torch.profiler`` (CPU + CUDA activities) and writes a Chrome trace.
"""
from __future__ import annotations

import contextlib
import os

import torch


class RoundTimer:
    def __init__(self):
        self.enabled = torch.cuda.is_available()
        self._events = []           # (round, event)
        self._last = None

    def mark(self, round_idx: int):
        if not self.enabled:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._events.append((round_idx, ev))

    def drain(self):
        """[(round, device ms since the previous mark)] for every mark whose event has completed."""
        out = []
        while self._events and self._events[0][1].query():
            r, ev = self._events.pop(0)
            if self._last is not None:
                out.append((r, self._last.elapsed_time(ev)))
            self._last = ev
        return out


_NVTX = os.environ.get("FLUTE_NVTX") == "1"


@contextlib.contextmanager
def nvtx_range(name: str):
    if _NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class TraceWindow:
    """Round-indexed torch.profiler window driven by ``step(round_idx)`` at the start of every round."""

    def __init__(self, spec=None):
        spec = spec if spec is not None else os.environ.get("FLUTE_TRACE", "")
        self.path, self.first, self.count, self._prof = None, 0, 0, None
        if spec:
            parts = spec.split(":")
            self.path = parts[0]
            self.first = int(parts[1]) if len(parts) > 1 else 3
            self.count = int(parts[2]) if len(parts) > 2 else 2

    def step(self, round_idx: int):
        if self.path is None:
            return
        if self._prof is None and round_idx == self.first:
            from torch.profiler import ProfilerActivity, profile
            acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
            self._prof = profile(activities=acts)
            self._prof.__enter__()
        elif self._prof is not None and round_idx >= self.first + self.count:
            self.close()

    def close(self):
        if self._prof is not None:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self._prof.__exit__(None, None, None)
            self._prof.export_chrome_trace(self.path)
            self._prof, self.path = None, None


class PhaseTimer:
    """Device time per named phase of a round (``bench.py``: the BASELINE metric asks for the *exposed* — not overlapped
    with client compute — broadcast + gather time per round).  ``with PHASES.phase("gather"):`` brackets the enqueued
    work with two CUDA events on the current stream; nothing is synchronised until :meth:`totals`.  Disabled (zero
    cost) unless ``enable()`` was called.  Everything of a round runs on one stream, so a phase's event-to-event time
    IS its exposed time."""

    def __init__(self):
        self.enabled = False
        self._spans = []            # (name, start_event, end_event)

    def enable(self, on=True):
        """Switching ON starts a fresh measurement; switching OFF keeps the recorded spans for :meth:`totals`."""
        on = bool(on) and torch.cuda.is_available()
        if on:
            self._spans = []
        self.enabled = on

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        try:
            yield
        finally:
            b.record()
            self._spans.append((name, a, b))

    def totals(self, reset=True):
        """{phase: total device ms} of every span recorded so far (synchronises)."""
        if not self._spans:
            return {}
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self._spans:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        if reset:
            self._spans = []
        return out


PHASES = PhaseTimer()
