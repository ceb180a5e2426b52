"""nlg_gru task model (next-word prediction on Reddit)."""
from msrflute_b200.models.gru_lm import GRU, GRU2, Embedding  # noqa: F401
