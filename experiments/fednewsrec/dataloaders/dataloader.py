"""Training batches stack (click, candidates) pairs; an evaluation batch is ONE impression (variable number of
candidates) — ref. ``experiments/fednewsrec/dataloaders/dataloader.py:37-41``."""
import numpy as np
import torch

from msrflute_b200.core.dataloader import BaseDataLoader
from experiments.fednewsrec.dataloaders.dataset import Dataset


class DataLoader(BaseDataLoader):
    def __init__(self, mode, num_workers=0, **kwargs):
        args = kwargs["args"]
        self.batch_size = args["batch_size"] if mode == "train" else 1
        dataset = Dataset(data=kwargs["data"], test_only=(mode != "train"), user_idx=kwargs.get("user_idx", None), args=args)
        super().__init__(dataset, batch_size=self.batch_size, shuffle=(mode == "train"), num_workers=num_workers,
                         collate_fn=self.collate_train if mode == "train" else self.collate_eval)

    @staticmethod
    def collate_train(batch):
        xs, ys = zip(*batch)
        click = torch.as_tensor(np.stack([x[0] for x in xs]))
        cand = torch.as_tensor(np.stack([x[1] for x in xs]))
        return {"x": (click, cand), "y": torch.as_tensor(np.asarray(ys))}

    @staticmethod
    def collate_eval(batch):
        (click, cand), y = batch[0]
        return {"x": (torch.as_tensor(click), torch.as_tensor(cand)), "y": torch.as_tensor(np.asarray(y))}
