"""Model-specific config class looked up by ``ModelConfig.from_dict`` (``<model_type>Config`` next to the model
file — the reference's plug-in hook, ``core/config.py:100-116``; ref. ``experiments/nlg_gru/config.py``)."""
from msrflute_b200.core.config import ModelConfig


class GRUConfig(ModelConfig):
    """GRU LM hyper-parameters: ``embed_dim``, ``vocab_size``, ``hidden_dim``, ``OOV_correct``, ``weight_init``."""
    DEFAULTS = dict(embed_dim=160, vocab_size=10000, hidden_dim=512, OOV_correct=False, weight_init="default")

    @staticmethod
    def from_dict(config):
        out = GRUConfig(GRUConfig.DEFAULTS)
        out.update(config)
        return out
