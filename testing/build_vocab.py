"""Build a word vocabulary from LEAF-style JSON user blobs (``--data-dir`` with ``*.json`` / ``*.txt`` files).

Counts whitespace tokens of every sample of every user, keeps the ``--vocab-size`` most frequent words (ties broken
alphabetically so the output is deterministic) behind the special tokens, and writes ``vocab_reddit.vocab`` to
``--target-dir`` as JSON ``{'vocab': {word: id}, 'size': n, 'unk_symbol': id, 'pad_symbol': id}`` — the structure the
reference's ``testing/build_vocab.py`` pickles (JSON here: readable and safe to load)."""
import argparse
import collections
import json
import os

SPECIALS = ["<PAD>", "<UNK>", "<BOS>", "<EOS>"]


def iter_samples(blob):
    for entry in blob["user_data"].values():
        xs = entry["x"] if isinstance(entry, dict) and "x" in entry else entry
        for s in xs:
            if isinstance(s, str):
                yield s.split()
            else:                                   # nested token lists (raw LEAF reddit)
                for piece in s:
                    yield piece.split() if isinstance(piece, str) else [str(t) for t in piece]


def count_words(paths, counter=None):
    counter = counter if counter is not None else collections.Counter()
    for p in paths:
        with open(p, encoding="utf8") as f:
            blob = json.load(f)
        for toks in iter_samples(blob):
            counter.update(t for t in toks if t not in SPECIALS)
    return counter


def build_vocab(counter, vocab_size=10000):
    words = [w for w, _ in sorted(counter.items(), key=lambda kv: (-kv[1], kv[0]))[:vocab_size]]
    vocab = {w: i for i, w in enumerate(SPECIALS + words)}
    return {"vocab": vocab, "size": len(vocab), "unk_symbol": vocab["<UNK>"], "pad_symbol": vocab["<PAD>"]}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--vocab-size", type=int, default=10000)
    ap.add_argument("--target-dir", default="./models")
    a = ap.parse_args(argv)
    files = sorted(os.path.join(a.data_dir, f) for f in os.listdir(a.data_dir) if f.endswith((".json", ".txt")))
    vocab = build_vocab(count_words(files), a.vocab_size)
    os.makedirs(a.target_dir, exist_ok=True)
    out = os.path.join(a.target_dir, "vocab_reddit.vocab")
    with open(out, "w", encoding="utf8") as f:
        json.dump(vocab, f)
    print("wrote {} ({} words)".format(out, vocab["size"]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
