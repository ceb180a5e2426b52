"""Loader for the in-tree CUDA extension (``msrflute_b200/_C*.so``).

The extension is built IN-TREE by ``__graft_entry__.build()`` /
``python setup.py build_ext --inplace`` for ``sm_100a`` only.  Policy:

* on a machine without CUDA (CI, gloo/CPU jobs) the ops fall back to their
  PyTorch reference implementations — the same code the tests use as oracle;
* on a machine WITH a GPU a missing extension is an error, loudly, unless
  ``FLUTE_ALLOW_FALLBACK=1`` — a silent eager fallback on a B200 would make the
  benchmarks meaningless.
"""
import importlib
import os

import torch

_EXT = None
_TRIED = False


def load(required: bool = False):
    global _EXT, _TRIED
    if _EXT is not None:
        return _EXT
    if not _TRIED:
        _TRIED = True
        try:
            _EXT = importlib.import_module("msrflute_b200._C")
        except ImportError as e:  # not built
            _EXT = None
            globals()["_ERR"] = e
    if _EXT is None and (required or (torch.cuda.is_available() and os.environ.get("FLUTE_ALLOW_FALLBACK") != "1")):
        raise RuntimeError(
            "msrflute_b200._C (sm_100a CUDA extension) is not built: run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python setup.py build_ext --inplace`. Set FLUTE_ALLOW_FALLBACK=1 to run the slow "
            "PyTorch reference path on purpose. Import error: {}".format(globals().get("_ERR")))
    return _EXT


def available() -> bool:
    try:
        return load(required=False) is not None
    except RuntimeError:
        return False


def use_cuda_kernels(*tensors) -> bool:
    """True when every tensor is on a CUDA device and the extension is loaded."""
    if not tensors or not all(t.is_cuda for t in tensors if t is not None):
        return False
    return load() is not None


#: counts launches of OUR kernels (bench.py reports it as ``gpu_launches``)
LAUNCH_COUNTER = {"n": 0}


def count_launch(n: int = 1):
    LAUNCH_COUNTER["n"] += n
