"""LR-MNIST task model (BASELINE config #1).  Class name / config keys as in the reference plug-in."""
from msrflute_b200.models.lr import LR, LogisticRegression  # noqa: F401
