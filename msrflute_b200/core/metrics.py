"""Evaluation metric aggregation (ref. ``core/metrics.py:19-79``).

Sample-weighted mean of every metric a model's ``inference`` returns, plus the
loss.  Unlike the reference (``metrics.py:37-38``: ``model.loss(batch).item()``
then ``model.inference(batch)`` — two forwards and a host sync per batch) this
uses ``model.loss_and_metrics`` (one forward when the model provides it) and
keeps the running sums on the device; a single D2H happens at the end.
"""
import logging

import numpy as np
import torch

from ..utils import print_rank


def _as_metric(v):
    if isinstance(v, dict):
        return v["value"], bool(v.get("higher_is_better", True))
    return v, True


class Metrics:
    def compute_metrics(self, dataloader, model):
        print_rank("Computing metrics", logging.DEBUG)
        return self.call_inference(dataloader, model)

    def call_inference(self, dataloader, model):
        sums, hib = {}, {}
        outs = {"probabilities": [], "predictions": [], "labels": []}
        counter = 0
        fused = hasattr(model, "loss_and_metrics")
        with torch.no_grad():
            for batch in dataloader:
                if fused:
                    loss, res = model.loss_and_metrics(batch)
                else:
                    loss, res = model.loss(batch), model.inference(batch)
                res = dict(res)
                output = res.pop("output", None)
                bs = res.pop("batch_size")
                res["loss"] = {"value": loss.detach() if torch.is_tensor(loss) else loss, "higher_is_better": False}
                for k, v in res.items():
                    val, h = _as_metric(v)
                    hib[k] = h
                    val = val.detach().float() if torch.is_tensor(val) else float(val)
                    sums[k] = sums.get(k, 0.0) + val * bs
                if isinstance(output, dict):
                    for k in outs:
                        outs[k].append(output[k])
                counter += bs
        for k in outs:
            outs[k] = np.concatenate(outs[k]) if outs[k] else []
        model.set_train()
        metrics = {}
        for k, s in sums.items():
            s = s.item() if torch.is_tensor(s) else s
            metrics[k] = {"value": s / max(counter, 1), "higher_is_better": hib[k]}
        print_rank(f"validation examples {counter}", loglevel=logging.DEBUG)
        return outs, metrics
