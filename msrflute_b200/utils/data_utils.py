"""Batch samplers (ref. ``utils/data_utils.py``: ``BatchSampler`` :9-39,
``DynamicBatchSampler`` :42-120)."""
import logging
import random

from torch.utils.data import sampler

from .utils import AverageMeter


class BatchSampler(sampler.Sampler):
    """Contiguous index blocks visited in random order, so neighbouring
    (similar-length) samples share a batch and padding stays small."""

    def __init__(self, dataset, batch_size, randomize=True, drop_last=False):
        self.dataset, self.batch_size, self.randomize = dataset, batch_size, randomize
        n = len(dataset)
        self.batches = [range(b, min(b + batch_size, n)) for b in range(0, n, batch_size)]
        if drop_last and self.batches and len(self.batches[-1]) < batch_size:
            self.batches.pop()

    def __iter__(self):
        if self.randomize:
            random.shuffle(self.batches)
        return iter(self.batches)

    def __len__(self):
        return len(self.batches) * self.batch_size


class DynamicBatchSampler(sampler.Sampler):
    """Variable batch size under a frame budget.

    Utterances (``dataset.utt_list[i]['duration']``) are sorted by duration and
    packed greedily while ``sum(frames) <= frames_threshold`` (and
    ``len <= max_batch_size`` when non-zero); with ``unsorted_batch`` the order
    is kept and only ``max_batch_size`` applies.  Batches are shuffled per epoch.
    """

    def __init__(self, sampler, frames_threshold, max_batch_size=0, unsorted_batch=False, fps=1000 / 30):
        self.sampler, self.frames_threshold = sampler, frames_threshold
        self.max_batch_size, self.unsorted_batch = max_batch_size, unsorted_batch
        utts = self.sampler.dataset.utt_list
        items = [(i, utts[i]["duration"]) for i in self.sampler]
        if not unsorted_batch:
            items.sort(key=lambda e: e[1])
        meter = AverageMeter("Padding Efficiency")
        batches, cur, cur_frames, cur_max = [], [], 0, 0

        def flush():
            if cur and cur_max > 0:
                meter.add(cur_frames, cur_max * len(cur))
                batches.append(list(cur))

        for idx, dur in items:
            if dur <= 0:
                continue
            frames = dur * fps
            if unsorted_batch:
                fits = len(cur) < max_batch_size
            else:
                fits = cur_frames + frames <= frames_threshold and (max_batch_size == 0 or len(cur) < max_batch_size)
            if fits:
                cur.append(idx)
                cur_frames += frames
                cur_max = max(cur_max, frames)
            else:
                flush()
                # the reference starts the new batch EMPTY but pre-charged with this
                # utterance's frames (data_utils.py:101-104), silently dropping it;
                # we keep the utterance.
                cur, cur_frames, cur_max = [idx], frames, frames
        flush()
        self.batches = batches
        if batches:
            meter.display_results(loglevel=logging.DEBUG)

    def __iter__(self):
        random.shuffle(self.batches)
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)
