"""MIND dataset reader + negative sampler (ref. ``experiments/fednewsrec/dataloaders/preprocess_mind.py``).

``read_news`` → ``news.tsv`` (doc id, category, sub-category, title) to dictionaries; ``get_doc_input`` → title word-id
matrix (30 words); ``read_clickhistory`` / ``parse_user`` → per-user click history (last 50) + impressions;
``get_train_input`` → (1 positive + ``npratio`` sampled negatives) candidate sets with the positive's position as label;
``get_test_input`` → one ranking instance per impression.  Tokenisation is ``nltk.word_tokenize`` when nltk is
installed, else a regex splitter.  ``synthetic_mind`` fabricates a small corpus with the same structures.
"""
import os
import random
import re

import numpy as np

MAX_TITLE, MAX_HIST, NPRATIO = 30, 50, 4


def tokenize(text):
    try:
        from nltk.tokenize import word_tokenize
        return word_tokenize(text)
    except Exception:
        return re.findall(r"\w+|[^\w\s]", text)


def read_news(root, modes):
    news, news_index, word_dict, cats, subcats = {}, {}, {}, [], []
    for mode in modes:
        with open(os.path.join(root, mode, "news.tsv"), encoding="utf8") as f:
            for line in f:
                doc_id, vert, subvert, title = line.strip("\n").split("\t")[:4]
                if doc_id in news_index:
                    continue
                news_index[doc_id] = len(news_index) + 1
                cats.append(vert); subcats.append(subvert)
                words = tokenize(title.lower())
                news[doc_id] = [vert, subvert, words]
                for w in words:
                    word_dict.setdefault(w.lower(), len(word_dict) + 1)
    cat_d = {c: i + 1 for i, c in enumerate(sorted(set(cats)))}
    sub_d = {c: i + 1 for i, c in enumerate(sorted(set(subcats)))}
    return news, news_index, cat_d, sub_d, word_dict


def load_matrix(embedding_path, word_dict, dim=300):
    mat = np.zeros((len(word_dict) + 1, dim), dtype=np.float32)
    have = []
    with open(os.path.join(embedding_path, "glove.840B.300d.txt"), "rb") as f:
        for line in f:
            parts = line.split()
            if len(parts) != dim + 1:
                continue
            w = parts[0].decode()
            if w in word_dict:
                mat[word_dict[w]] = np.asarray(parts[1:], dtype=np.float32)
                have.append(w)
    return mat, have


def get_doc_input(news, news_index, word_dict):
    out = np.zeros((len(news) + 1, MAX_TITLE), dtype=np.int64)
    for doc_id, (_, _, title) in news.items():
        ids = [word_dict.get(w.lower(), 0) for w in title[:MAX_TITLE]]
        out[news_index[doc_id], :len(ids)] = ids
    return out


def read_clickhistory(root, mode, news_index):
    sessions = []
    with open(os.path.join(root, mode, "behaviors.tsv"), encoding="utf8") as f:
        for line in f:
            _, uid, _, click, imps = line.strip("\n").split("\t")
            clicks = [c for c in click.split() if c in news_index]
            pos, neg = [], []
            for imp in imps.split():
                doc, lab = imp.rsplit("-", 1)
                (pos if lab == "1" else neg).append(doc)
            sessions.append([uid, clicks, pos, neg])
    return sessions


def parse_user(news_index, sessions):
    click = np.zeros((len(sessions), MAX_HIST), dtype=np.int64)
    for i, (_, clicks, _, _) in enumerate(sessions):
        ids = [news_index[c] for c in clicks][-MAX_HIST:]
        if ids:
            click[i, -len(ids):] = ids
    return {"click": click}


def newsample(pool, ratio):
    if ratio > len(pool):
        return random.sample(pool * (ratio // max(len(pool), 1) + 1), ratio) if pool else [0] * ratio
    return random.sample(pool, ratio)


def get_train_input(sessions, news_index, npratio=NPRATIO):
    cands, labels, users = [], [], []
    for si, (_, _, pos, neg) in enumerate(sessions):
        neg_ids = [news_index[n] for n in neg if n in news_index]
        for p in pos:
            if p not in news_index:
                continue
            docs = newsample(neg_ids, npratio) + [news_index[p]]
            order = list(range(npratio + 1))
            random.shuffle(order)
            cands.append([docs[i] for i in order])
            labels.append(order.index(npratio))
            users.append(si)
    return np.asarray(cands, dtype=np.int64), np.asarray(users, dtype=np.int64), np.asarray(labels, dtype=np.int64)


def get_test_input(sessions, news_index):
    imps = []
    for si, (_, _, pos, neg) in enumerate(sessions):
        docs = [news_index[d] for d in pos + neg if d in news_index]
        labs = [1] * len([d for d in pos if d in news_index]) + [0] * len([d for d in neg if d in news_index])
        if docs and 0 < sum(labs) < len(labs):
            imps.append({"user": si, "docs": np.asarray(docs, dtype=np.int64), "labels": np.asarray(labs, dtype=np.int64)})
    return imps


class MIND:
    """In-memory MIND split: ``news_title`` [n_news+1, 30], per-session click history, train candidates / test impressions."""

    def __init__(self, root=None, mode="train", n_news=400, n_users=40, vocab=5000, seed=0):
        if root and os.path.exists(os.path.join(root, mode, "news.tsv")):
            news, self.news_index, _, _, self.word_dict = read_news(root, ["train", "val"])
            self.news_title = get_doc_input(news, self.news_index, self.word_dict)
            self.sessions = read_clickhistory(root, mode, self.news_index)
        else:
            self.news_title, self.sessions, self.news_index = synthetic_mind(n_news, n_users, vocab, seed)
        self.user = parse_user(self.news_index, self.sessions)


def synthetic_mind(n_news=400, n_users=40, vocab=5000, seed=0):
    rng = np.random.default_rng(seed)
    titles = np.zeros((n_news + 1, MAX_TITLE), dtype=np.int64)
    topic = rng.integers(0, 8, size=n_news + 1)
    for i in range(1, n_news + 1):
        L = int(rng.integers(5, MAX_TITLE))
        titles[i, :L] = 1 + topic[i] * (vocab // 8) + rng.integers(0, vocab // 8, size=L)
    index = {"N{}".format(i): i for i in range(1, n_news + 1)}
    sessions = []
    for u in range(n_users):
        fav = int(rng.integers(0, 8))
        liked = [i for i in range(1, n_news + 1) if topic[i] == fav]
        other = [i for i in range(1, n_news + 1) if topic[i] != fav]
        for _ in range(int(rng.integers(2, 5))):
            clicks = ["N{}".format(i) for i in rng.choice(liked, size=min(len(liked), int(rng.integers(5, 30))), replace=False)]
            pos = ["N{}".format(i) for i in rng.choice(liked, size=int(rng.integers(1, 3)), replace=False)]
            neg = ["N{}".format(i) for i in rng.choice(other, size=int(rng.integers(6, 15)), replace=False)]
            sessions.append(["U{}".format(u), clicks, pos, neg])
    return titles, sessions, index
