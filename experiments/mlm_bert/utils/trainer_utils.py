"""Small evaluation helpers (ref. ``experiments/mlm_bert/utils/trainer_utils.py``): ``set_seed``, ``EvalPrediction``,
``ComputeMetrics.compute_metrics`` = accuracy over the masked (label != −100) positions."""
import random
from typing import NamedTuple, Union, Tuple

import numpy as np
import torch


def set_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class EvalPrediction(NamedTuple):
    predictions: Union[np.ndarray, Tuple[np.ndarray]]
    label_ids: np.ndarray


class ComputeMetrics:
    @staticmethod
    def compute_metrics(p: EvalPrediction, mask=None):
        preds = torch.as_tensor(np.asarray(p.predictions))
        if preds.dim() == 3:
            preds = preds.argmax(dim=-1)
        labels = torch.as_tensor(np.asarray(p.label_ids))
        keep = labels != -100
        acc = ((preds == labels) & keep).sum().float() / keep.sum().clamp(min=1)
        return {"acc": acc}
