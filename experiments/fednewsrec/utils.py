"""Ranking metrics on numpy arrays (ref. ``experiments/fednewsrec/utils.py``); device versions live in
``msrflute_b200.models.newsrec``."""
import numpy as np


def dcg_score(y_true, y_score, k=10):
    order = np.argsort(y_score)[::-1]
    y = np.take(y_true, order[:k])
    return float(np.sum((2 ** y - 1) / np.log2(np.arange(len(y)) + 2)))


def ndcg_score(y_true, y_score, k=10):
    best = dcg_score(y_true, y_true, k)
    return dcg_score(y_true, y_score, k) / best if best > 0 else 0.0


def mrr_score(y_true, y_score):
    order = np.argsort(y_score)[::-1]
    y = np.take(y_true, order)
    return float(np.sum(y / (np.arange(len(y)) + 1)) / max(np.sum(y), 1e-12))
