"""Helpers shared by both benchmark arms (self-contained: imports nothing from msrflute_b200)."""
import argparse
import json
import os
import statistics
import subprocess
import threading
import time

HEADLINE_METRIC = "fl_rounds_per_sec_resnet18_fedcifar100"
BASELINE_PUBLISHED_ROUNDS_PER_SEC = 4000.0 / 6121.0     # RESNET_FEDCIFAR100, 4000 rounds in 01:42:01


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--task", default="cv_resnet_fedcifar100")
    p.add_argument("--norm", default="gn", choices=["gn", "bn"], help="ours only: GroupNorm (BASELINE.json) or "
                   "BatchNorm (what the reference actually instantiates)")
    p.add_argument("--comm", default="auto")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--clients-per-round", type=int, default=10)
    return p.parse_args(argv)


class ClockSampler:
    """Samples ``nvidia-smi`` SM clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_ms=200):
        self.gpu, self.period = gpu_index, period_ms
        self.proc, self.lines = None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", str(self.period)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
        except (OSError, FileNotFoundError):
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def emit(d):
    print(json.dumps(d), flush=True)
