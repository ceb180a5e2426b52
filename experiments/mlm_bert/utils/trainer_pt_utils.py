"""Tensor-nest helpers and length-grouped samplers (ref. ``experiments/mlm_bert/utils/trainer_pt_utils.py`` — HF Trainer
utilities the reference vendors).  Re-implemented compactly; ``LabelSmoother`` lives with the model."""
import math
from typing import Iterator, List, Optional

import numpy as np
import torch
from torch.utils.data import Sampler

from msrflute_b200.models.bert_mlm import LabelSmoother  # noqa: F401


def _map(fn, t):
    if isinstance(t, (list, tuple)):
        return type(t)(_map(fn, x) for x in t)
    if isinstance(t, dict):
        return {k: _map(fn, v) for k, v in t.items()}
    return fn(t)


def nested_detach(tensors):
    return _map(lambda x: x.detach(), tensors)


def nested_numpify(tensors):
    return _map(lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x), tensors)


def _pad_cat(a, b, padding_index=-100):
    """Concatenate on dim 0, right-padding dim 1 to the longer of the two."""
    is_np = isinstance(a, np.ndarray)
    if is_np:
        a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dim() == 1 or a.shape[1] == b.shape[1]:
        out = torch.cat((a, b), dim=0)
    else:
        L = max(a.shape[1], b.shape[1])
        out = a.new_full((a.shape[0] + b.shape[0], L) + tuple(a.shape[2:]), padding_index)
        out[:a.shape[0], :a.shape[1]] = a
        out[a.shape[0]:, :b.shape[1]] = b
    return out.numpy() if is_np else out


def nested_concat(tensors, new_tensors, padding_index=-100):
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(nested_concat(t, n, padding_index) for t, n in zip(tensors, new_tensors))
    return _pad_cat(tensors, new_tensors, padding_index)


class DistributedTensorGatherer:
    """Collects per-process evaluation arrays in dataset order (round-robin sharded samplers) and truncates the
    padding samples at ``finalize``."""

    def __init__(self, world_size, num_samples, make_multiple_of=None, padding_index=-100):
        self.world_size, self.num_samples = world_size, num_samples
        total = world_size if make_multiple_of is None else world_size * make_multiple_of
        self.total_samples = int(np.ceil(num_samples / total)) * total
        self.process_length = self.total_samples // world_size
        self._storage, self._offsets, self.padding_index = None, None, padding_index

    def add_arrays(self, arrays, want_masked=False):
        if arrays is None:
            return
        arrays = np.asarray(arrays)
        if self._storage is None:
            self._storage = np.full((self.total_samples,) + arrays.shape[1:], self.padding_index, dtype=arrays.dtype)
            self._offsets = list(range(0, self.total_samples, self.process_length))
        chunk = arrays.shape[0] // self.world_size
        for i in range(self.world_size):
            part = arrays[i * chunk:(i + 1) * chunk]
            if part.ndim > 1 and part.shape[1] > self._storage.shape[1]:
                grown = np.full((self.total_samples, part.shape[1]) + self._storage.shape[2:], self.padding_index,
                                dtype=self._storage.dtype)
                grown[:, :self._storage.shape[1]] = self._storage
                self._storage = grown
            sl = (slice(self._offsets[i], self._offsets[i] + chunk),) + ((slice(0, part.shape[1]),) if part.ndim > 1 else ())
            self._storage[sl] = part
            self._offsets[i] += chunk

    def finalize(self):
        return None if self._storage is None else self._storage[:self.num_samples]


def get_length_grouped_indices(lengths, batch_size, mega_batch_mult=None, generator=None):
    """Shuffle, cut into mega-batches, sort each by length (descending) so batches need little padding; the longest
    sample is moved to the very first batch so an OOM shows up immediately."""
    if mega_batch_mult is None:
        mega_batch_mult = max(min(len(lengths) // (batch_size * 4), 50), 1)
    idx = torch.randperm(len(lengths), generator=generator).tolist()
    mb = mega_batch_mult * batch_size
    megas = [sorted(idx[i:i + mb], key=lambda i: lengths[i], reverse=True) for i in range(0, len(lengths), mb)]
    firsts = [lengths[m[0]] for m in megas]
    k = int(np.argmax(firsts)) if firsts else 0
    if megas:
        megas[0][0], megas[k][0] = megas[k][0], megas[0][0]
    return [i for m in megas for i in m]


class LengthGroupedSampler(Sampler):
    def __init__(self, dataset, batch_size, lengths: Optional[List[int]] = None, model_input_name="input_ids"):
        self.dataset, self.batch_size = dataset, batch_size
        self.lengths = lengths if lengths is not None else [len(f[model_input_name]) for f in dataset]

    def __len__(self):
        return len(self.lengths)

    def __iter__(self):
        return iter(get_length_grouped_indices(self.lengths, self.batch_size))


class DistributedLengthGroupedSampler(Sampler):
    def __init__(self, dataset, batch_size, num_replicas=1, rank=0, seed=0, drop_last=False, lengths=None,
                 model_input_name="input_ids"):
        self.dataset, self.batch_size, self.num_replicas, self.rank = dataset, batch_size, num_replicas, rank
        self.epoch, self.seed, self.drop_last = 0, seed, drop_last
        self.lengths = lengths if lengths is not None else [len(f[model_input_name]) for f in dataset]
        n = len(self.lengths)
        self.num_samples = n // num_replicas if drop_last else math.ceil(n / num_replicas)
        self.total_size = self.num_samples * num_replicas

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self.num_samples

    def __iter__(self) -> Iterator:
        g = torch.Generator()
        g.manual_seed(self.seed + self.epoch)
        idx = get_length_grouped_indices(self.lengths, self.batch_size, generator=g)
        idx = idx[:self.total_size] if self.drop_last else idx + idx[:self.total_size - len(idx)]
        return iter(idx[self.rank:self.total_size:self.num_replicas])
