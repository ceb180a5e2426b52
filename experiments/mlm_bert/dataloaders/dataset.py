"""Per-user text → fixed-length MLM frames (ref. ``experiments/mlm_bert/dataloaders/dataset.py``).

``user_data[u]`` is a list of lines (each a list of sentences, as in the reference's Reddit json).  Sentences with
fewer than ``min_words_per_utt`` words are dropped, at most ``max_samples_per_user`` kept; the user's sentences are
tokenised, concatenated (padding the tail when ``padding``) and cut into ``max_seq_length`` frames
(``group_texts``, ref :151-187); ``num_samples`` is updated to the number of frames.  ``process_line_by_line`` keeps
one padded sequence per sentence instead.

Tokeniser: a HuggingFace tokenizer if ``model_name_or_path`` / ``tokenizer_name`` resolves to LOCAL files, otherwise
``HashTokenizer`` (deterministic word → id hashing with [CLS]/[SEP]/[PAD]/[MASK] ids) — there is no network here.
With ``data=None`` a synthetic corpus is generated.
"""
import hashlib
import itertools
import json
import logging
import os

from msrflute_b200.core.dataset import BaseDataset
from msrflute_b200.data import synthetic
from msrflute_b200.utils import print_rank


class HashTokenizer:
    pad_token_id, cls_token_id, sep_token_id, mask_token_id, unk_token_id = 0, 1, 2, 3, 4
    n_special = 5

    def __init__(self, vocab_size=30522, model_max_length=512):
        self.vocab_size, self.model_max_length = vocab_size, model_max_length

    def __len__(self):
        return self.vocab_size

    def _id(self, word):
        h = int(hashlib.blake2s(word.lower().encode("utf8"), digest_size=4).hexdigest(), 16)
        return self.n_special + h % (self.vocab_size - self.n_special)

    def __call__(self, lines, truncation=True, max_length=512, padding=False, return_special_tokens_mask=True):
        out = {"input_ids": [], "attention_mask": [], "special_tokens_mask": []}
        for line in lines:
            ids = [self.cls_token_id] + [self._id(w) for w in line.split()][:max_length - 2 if truncation else None] + [self.sep_token_id]
            stm = [1] + [0] * (len(ids) - 2) + [1]
            att = [1] * len(ids)
            if padding == "max_length":
                pad = max_length - len(ids)
                ids, stm, att = ids + [self.pad_token_id] * pad, stm + [1] * pad, att + [0] * pad
            out["input_ids"].append(ids); out["attention_mask"].append(att); out["special_tokens_mask"].append(stm)
        return out


def load_tokenizer(args):
    name = args.get("tokenizer_name", args.get("model_name_or_path", None))
    if isinstance(name, str) and os.path.isdir(name):
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(name, local_files_only=True, use_fast=args.get("tokenizer_type_fast", False))
    return HashTokenizer(int(args.get("vocab_size", 30522)), int(args.get("model_max_length", 512)))


class Dataset(BaseDataset):
    def __init__(self, data, args, tokenizer=None, test_only=False, user_idx=0, max_samples_per_user=-1,
                 min_words_per_utt=5, **kwargs):
        self.utt_list, self.test_only = [], test_only
        self.padding = args.get("padding", True)
        self.max_seq_length = args.get("max_seq_length", 256)
        self.max_samples_per_user, self.min_num_words = max_samples_per_user, min_words_per_utt
        self.process_line_by_line = args.get("process_line_by_line", False)
        self.tokenizer = tokenizer or load_tokenizer(args)
        self.max_seq_length = min(self.max_seq_length or 512, getattr(self.tokenizer, "model_max_length", 512), 512)
        if data is None:
            st = synthetic.make_token_lists(num_users=64 if not test_only else 8, mean_samples=12, max_len=24,
                                            vocab=2000, seed=51 if not test_only else 52, as_text=True)
            data = {"users": st["users"], "num_samples": st["num_samples"],
                    "user_data": {u: [[s] for s in v] for u, v in st["user_data"].items()}}
        self.load_data(data, user_idx)
        if user_idx != -1 and not self.process_line_by_line:
            self.post_process_list()

    def __len__(self):
        return len(self.utt_list)

    def __getitem__(self, idx):
        if self.process_line_by_line:
            enc = self.tokenizer([self.utt_list[idx]["src_text"]], truncation=True, max_length=self.max_seq_length,
                                 padding="max_length", return_special_tokens_mask=True)
            return {k: v[0] for k, v in enc.items()}
        return self.utt_list[idx]

    def load_data(self, orig_strct, user_idx):
        if isinstance(orig_strct, str):
            with open(orig_strct, "r") as f:
                orig_strct = json.load(f)
        self.user_list, self.num_samples = orig_strct["users"], list(orig_strct["num_samples"])
        self.user_data = orig_strct["user_data"]
        if user_idx == -1:
            return
        if self.test_only:
            self.user = "test_only"
            for i, u in enumerate(self.user_list):
                self.num_samples[i] = self.process_user(u, self.user_data[u])
        else:
            self.user = self.user_list[user_idx]
            self.num_samples[user_idx] = self.process_user(self.user, self.user_data[self.user])
        if not self.utt_list:
            self.utt_list = [{"src_text": "N/A", "duration": 0, "loss_weight": 1.0}]

    def process_user(self, user, user_data):
        counter = 0
        lines = user_data["x"] if isinstance(user_data, dict) else user_data
        for line in lines:
            for e in ([line] if isinstance(line, str) else line):
                if len(e.split()) < self.min_num_words:
                    continue
                if -1 < self.max_samples_per_user <= counter:
                    return counter
                counter += 1
                self.utt_list.append({"src_text": e, "duration": len(e.split()), "loss_weight": 1.0})
        return counter

    def post_process_list(self):
        lines = [u["src_text"] for u in self.utt_list if u["src_text"] != "N/A" and u["src_text"].strip()]
        if not lines:
            self.utt_list = [{"input_ids": [0, 2], "special_tokens_mask": [1, 1], "attention_mask": [0, 0]}]
            return
        enc = dict(self.tokenizer(lines, truncation=True, max_length=512, padding=False, return_special_tokens_mask=True))
        enc = {k: [list(x) for x in v] for k, v in enc.items()}
        L = self.max_seq_length
        if self.padding:
            total = sum(len(x) for x in enc["input_ids"])
            pad = L - (total % L)
            fill = {"input_ids": getattr(self.tokenizer, "pad_token_id", 0) or 0, "attention_mask": 0,
                    "special_tokens_mask": 1, "token_type_ids": 0}
            for k in enc:
                enc[k].append([fill.get(k, 0)] * pad)
        cat = {k: list(itertools.chain.from_iterable(v)) for k, v in enc.items()}
        total = (len(cat["input_ids"]) // L) * L
        self.utt_list = [{k: t[i:i + L] for k, t in cat.items()} for i in range(0, total, L)]
        print_rank("Finished reshaping in sequences of {} frames".format(len(self.utt_list)), logging.DEBUG)
        if not self.test_only:
            self.num_samples[self.user_list.index(self.user)] = len(self.utt_list)
        if not self.utt_list:
            self.utt_list = [{"input_ids": [0, 2], "special_tokens_mask": [1, 1], "attention_mask": [0, 0]}]
