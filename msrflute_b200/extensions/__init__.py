"""Optional algorithmic extensions: privacy (DP), gradient quantization, RL aggregation weights."""
from .quantization import quant_model  # noqa: F401
from . import privacy  # noqa: F401


def __getattr__(name):
    if name == "RL":
        from .RL import RL
        return RL
    raise AttributeError(name)
