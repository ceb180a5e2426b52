"""Offline stand-in for ``azureml.core`` (not installable here; the reference imports it at module import time).

``Run.get_context()`` returns a process-wide offline run whose ``log`` keeps the history and fires optional
callbacks — ``baseline/run_reference.py`` uses one to timestamp the reference's own ``secsPerRoundTotal``
log line, i.e. round boundaries are observed without touching reference code."""
import os

_RUN = None


class Run:
    def __init__(self):
        # experiment_name = "-".join(id.split("-")[-4:-2])  (reference e2e_trainer.py:222-223)
        self.id = os.environ.get("FLUTE_REF_RUN_ID", "offline-ref-bench-0-0")
        self.history = {}
        self.callbacks = []
        self.input_datasets = {"input": os.environ.get("FLUTE_REF_DATA", ".")}

    @staticmethod
    def get_context():
        global _RUN
        if _RUN is None:
            _RUN = Run()
        return _RUN

    def log(self, key, value):
        self.history.setdefault(key, []).append(value)
        for cb in self.callbacks:
            cb(key, value)

    def log_list(self, key, value):
        self.log(key, value)
