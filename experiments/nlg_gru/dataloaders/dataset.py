"""Reddit-style text dataset (ref. ``experiments/nlg_gru/dataloaders/dataset.py``): ``user_data[u]['x']`` is a list
of utterances (strings, word lists, or pre-encoded id lists); utterances with ≤ ``min_num_words`` tokens are
dropped, longer than ``max_num_words`` truncated; ``utt_list[i]['duration']`` feeds the dynamic batch sampler.
With ``data=None`` a synthetic Zipf corpus is generated (pre-encoded)."""
import json
import logging

import numpy as np

from msrflute_b200.core.dataset import BaseDataset
from msrflute_b200.data import synthetic
from msrflute_b200.utils import print_rank
from experiments.nlg_gru.utils.utility import case_backoff_batch, load_vocab, make_vocab, to_indices


class Dataset(BaseDataset):
    def __init__(self, data, min_num_words=2, max_num_words=25, test_only=False, user_idx=0, vocab_dict=None,
                 preencoded=False, **kwargs):
        args = kwargs.get("args", {}) or {}
        self.utt_list = []
        self.test_only = test_only
        self.max_num_words = args.get("max_num_words", max_num_words) or max_num_words
        self.min_num_words = min_num_words
        self.preencoded = args.get("preencoded", preencoded)
        vocab_path = args.get("vocab_dict", vocab_dict)
        if data is None:
            vsize = int(args.get("vocab_size", 10000))
            data = synthetic.make_token_lists(num_users=100 if not test_only else 20, vocab=vsize,
                                              max_len=self.max_num_words, seed=11 if not test_only else 12)
            data["user_data"] = {u: {"x": v} for u, v in data["user_data"].items()}
            self.preencoded = True
            self.vocab = make_vocab(["w{}".format(i) for i in range(vsize)])
        else:
            self.vocab = load_vocab(vocab_path) if vocab_path else make_vocab([])
        self.vocab_size = len(self.vocab.idx_to_term)
        self.load_data(data, user_idx)

    def __len__(self):
        return len(self.utt_list)

    def __getitem__(self, idx):
        if self.preencoded:
            batch = np.array([self.utt_list[idx]["src_text"]], dtype=np.int32)
        else:
            batch = to_indices(self.vocab, case_backoff_batch([self.utt_list[idx]["src_text"]], self.vocab.term_to_idx))
        return batch, self.user

    def load_data(self, orig_strct, user_idx):
        if isinstance(orig_strct, str):
            with open(orig_strct, "r") as f:
                orig_strct = json.load(f)
        self.user_list = orig_strct["users"]
        self.num_samples = orig_strct["num_samples"]
        self.user_data = orig_strct["user_data"]
        self.user = "test_only" if self.test_only else self.user_list[user_idx]
        if user_idx != -1:
            self.process_x(self.user_data)

    def process_x(self, user_data):
        print_rank("Processing data-structure: {} Utterances expected".format(sum(self.num_samples)), logging.DEBUG)
        for user in self.user_list:
            entries = user_data[user]["x"] if isinstance(user_data[user], dict) else user_data[user]
            for e in entries:
                words = e if isinstance(e, (list, tuple)) else e.split()
                if len(e) <= self.min_num_words:
                    continue
                words = list(words[:self.max_num_words])
                if words and isinstance(words[0], (int, np.integer)):
                    self.preencoded = True              # id lists need no vocabulary lookup
                self.utt_list.append({"src_text": words, "duration": len(words), "loss_weight": 1.0})
