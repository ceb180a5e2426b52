"""Compatibility alias for reference-style imports → ``msrflute_b200.utils``."""
import importlib as _il
_m = _il.import_module("msrflute_b200.utils")
_g = globals()
for _k in dir(_m):
    if not _k.startswith("__"):
        _g[_k] = getattr(_m, _k)
