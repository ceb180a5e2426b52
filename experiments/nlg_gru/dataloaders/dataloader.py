"""Dataloader for the text tasks (ref. ``experiments/nlg_gru/dataloaders/dataloader.py``): training uses the
frame-budget ``DynamicBatchSampler`` (budget = ``max_num_words``·… tokens, at most ``batch_size`` utterances),
evaluation fixed-size blocks; batches are padded with −1."""
import numpy as np
import torch
from torch.utils.data import RandomSampler, SequentialSampler

from msrflute_b200.core.dataloader import BaseDataLoader
from msrflute_b200.utils.data_utils import BatchSampler, DynamicBatchSampler
from experiments.nlg_gru.dataloaders.dataset import Dataset


class _Shuffled(RandomSampler):
    """RandomSampler that also exposes ``.dataset`` (what ``DynamicBatchSampler`` expects from its sampler)."""

    def __init__(self, dataset):
        super().__init__(dataset)
        self.dataset = dataset


class DataLoader(BaseDataLoader):
    def __init__(self, mode, num_workers=0, **kwargs):
        args = kwargs["args"]
        self.batch_size = args["batch_size"]
        dataset = Dataset(data=kwargs["data"], test_only=(mode != "train"), vocab_dict=args.get("vocab_dict"),
                          user_idx=kwargs["user_idx"], max_num_words=args.get("max_num_words", 25),
                          preencoded=args.get("preencoded", False), args=args)
        self.vocab_size = dataset.vocab_size
        if mode == "train":
            sampler = DynamicBatchSampler(_Shuffled(dataset), frames_threshold=args.get("max_num_words", 25) * self.batch_size
                                          if args.get("frames_budget_per_batch", False) else args.get("max_num_words", 25),
                                          max_batch_size=self.batch_size, unsorted_batch=args.get("unsorted_batch", False), fps=1)
        else:
            sampler = BatchSampler(dataset, batch_size=self.batch_size, randomize=False, drop_last=False)
        super().__init__(dataset, batch_sampler=sampler, num_workers=num_workers, collate_fn=self.collate_fn,
                         pin_memory=False)

    @staticmethod
    def collate_fn(batch):
        seqs, utt_ids = zip(*batch)
        lens = [len(s[0]) for s in seqs]
        out = np.full((len(seqs), max(lens)), -1, dtype=np.int64)
        for i, s in enumerate(seqs):
            out[i, :lens[i]] = np.squeeze(s, axis=0)
        return {"x": torch.from_numpy(out), "x_len": lens, "utt_ids": utt_ids, "total_frames": sum(lens),
                "total_frames_with_padding": int(np.prod(out.shape)), "loss_weight": None}
