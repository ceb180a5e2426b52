"""Vocabulary helpers for the text tasks (ref. ``experiments/nlg_gru/utils/utility.py``): vocab file loading,
case back-off (try the word, its lower-case and capitalised variants against the vocabulary) and padding to an
index array (OOV → 0, pad → −1)."""
from collections import namedtuple

import numpy as np

Vocab = namedtuple("Vocab", ["idx_to_term", "term_to_idx"])
_TR_LOWER = {ord("I"): "ı"}
_TR_UPPER = {ord("i"): "İ"}


def load_vocab(url):
    with open(url, "r", encoding="utf-8") as f:
        terms = [line.strip() for line in f]
    return Vocab(terms, {w: i for i, w in enumerate(terms)})


def make_vocab(terms):
    terms = list(terms)
    return Vocab(terms, {w: i for i, w in enumerate(terms)})


def to_indices(vocab, batch, ndim=2, oov_idx=0, pad_idx=-1):
    """Nested list of strings → int32 array; short rows padded with ``pad_idx``; unknown terms → ``oov_idx``
    (``None`` makes either situation an error)."""
    def look(term):
        return vocab.term_to_idx[term] if oov_idx is None else vocab.term_to_idx.get(term, oov_idx)

    if ndim == 1:
        return np.array([look(t) for t in batch], dtype=np.int32)
    if ndim == 2:
        length = max(len(r) for r in batch)
        if pad_idx is None and min(len(r) for r in batch) != length:
            raise ValueError("Padding required, but no pad_idx provided")
        return np.array([[look(t) for t in r] + [pad_idx] * (length - len(r)) for r in batch], dtype=np.int32)
    shape = [len(batch)]
    for _ in range(2, ndim):
        shape.append(len(batch[0]))
        batch = [x for sub in batch for x in sub]
    return to_indices(vocab, batch, 2, oov_idx, pad_idx).reshape(*shape, -1)


def _variants(word):
    yield word
    yield word.translate(_TR_LOWER).lower()
    yield word.lower()
    if len(word) > 1:
        yield word[0].translate(_TR_UPPER).upper() + word[1:]
        yield word[0].upper() + word[1:]
        yield word[0].translate(_TR_UPPER).upper() + word[1:].translate(_TR_LOWER).lower()
        yield word[0].upper() + word[1:].lower()
    else:
        yield word.translate(_TR_UPPER).upper()
        yield word.upper()


def case_backoff(word, vocab):
    if not isinstance(word, str):
        return word
    for v in _variants(word):
        if v in vocab:
            return v
    return word


def case_backoff_batch(batch, vocab):
    return [[case_backoff(w, vocab) for w in sent] for sent in batch]


def encode_data(data_dict, vocab):
    """{'users': …, 'user_data': {u: {'x': [[words]]}}} → same structure with int ids (pre-encoding)."""
    out = {"users": list(data_dict["users"]), "num_samples": list(data_dict["num_samples"]), "user_data": {}}
    for u in data_dict["users"]:
        sents = data_dict["user_data"][u]["x"] if isinstance(data_dict["user_data"][u], dict) else data_dict["user_data"][u]
        sents = [s.split() if isinstance(s, str) else s for s in sents]
        out["user_data"][u] = {"x": [to_indices(vocab, case_backoff_batch([s], vocab.term_to_idx))[0].tolist() for s in sents]}
    return out
