import numpy as np
import torch

from msrflute_b200.core.dataloader import BaseDataLoader
from experiments.semisupervision.dataloaders.dataset import Dataset


class DataLoader(BaseDataLoader):
    def __init__(self, mode, num_workers=0, **kwargs):
        args = kwargs["args"]
        self.batch_size = args["batch_size"]
        dataset = Dataset(data=kwargs["data"], test_only=(mode != "train"), user_idx=kwargs.get("user_idx", None), args=args)
        super().__init__(dataset, batch_size=self.batch_size, shuffle=(mode == "train"), num_workers=num_workers,
                         collate_fn=self.collate_fn)

    @staticmethod
    def collate_fn(batch):
        x, y = zip(*batch)
        return {"x": torch.as_tensor(np.stack(x)), "y": torch.as_tensor(np.asarray(y))}
