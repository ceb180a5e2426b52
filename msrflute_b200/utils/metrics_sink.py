"""Scalar-metric sink replacing ``azureml.core.Run`` (ref. imports it at module import time in
``e2e_trainer.py:21``, ``core/server.py:43-44`` …; offline it is a no-op/print).

``get_run()`` returns a process-wide ``Run`` whose ``log(key, value)`` appends to an in-memory history,
optionally mirrors to ``<log_dir>/metrics.jsonl`` and forwards to AzureML when that SDK is importable
and ``FLUTE_USE_AZUREML=1``.  ``Run.get_context()`` mirrors the AzureML call shape.
"""
import json
import os
import time
import uuid

_RUN = None


class Run:
    def __init__(self, run_id=None):
        self.id = run_id or os.environ.get("FLUTE_RUN_ID") or "OfflineRun_" + str(uuid.uuid4())
        self.history = {}
        self._fp = None
        self._aml = None
        if os.environ.get("FLUTE_USE_AZUREML") == "1":
            try:
                from azureml.core import Run as _AmlRun
                self._aml = _AmlRun.get_context()
                self.id = self._aml.id
            except Exception:
                self._aml = None
        self.input_datasets = {}

    @staticmethod
    def get_context():
        return get_run()

    def attach_file(self, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        self._fp = open(path, "a", buffering=1)

    def log(self, key, value):
        try:
            v = float(value)
        except (TypeError, ValueError):
            v = value
        self.history.setdefault(key, []).append(v)
        if self._fp is not None:
            self._fp.write(json.dumps({"t": time.time(), "k": key, "v": v if isinstance(v, (int, float, str, bool)) else str(v)}) + "\n")
        if self._aml is not None:
            self._aml.log(key, value)

    def last(self, key, default=None):
        h = self.history.get(key)
        return h[-1] if h else default


def get_run() -> Run:
    global _RUN
    if _RUN is None:
        _RUN = Run()
    return _RUN


def reset_run():
    global _RUN
    _RUN = None
