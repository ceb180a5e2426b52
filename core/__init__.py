"""Compatibility alias: FLUTE plug-ins written against the reference import ``core.*``; everything lives in ``msrflute_b200.core``."""
