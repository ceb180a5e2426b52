"""``BERTConfig`` — model-specific config class picked up by ``ModelConfig.from_dict`` (ref. ``experiments/mlm_bert/config.py``
and ``core/config.py:118-240``: ``BERTModelConfig`` / ``BERTTrainingConfig`` / ``BERTConfig``)."""
from msrflute_b200.core.config import BERTConfig as _Node, BERTModelConfig, BERTTrainingConfig, ModelConfig


class BERTConfig(ModelConfig):
    MODEL_DEFAULTS = dict(model_name=None, cache_dir=None, use_fast_tokenizer=False, mask_token="<mask>", task="mlm",
                          past_index=-1, prediction_loss_only=False, process_line_by_line=False)
    TRAINING_DEFAULTS = dict(seed=12345, label_smoothing_factor=0, batch_size=64, max_seq_length=256)

    @staticmethod
    def from_dict(config):
        out = BERTConfig(config)
        bert = _Node(config.get("BERT", {}) or {})
        model = BERTModelConfig(BERTConfig.MODEL_DEFAULTS); model.update(bert.get("model", {}) or {})
        training = BERTTrainingConfig(BERTConfig.TRAINING_DEFAULTS); training.update(bert.get("training", {}) or {})
        bert["model"], bert["training"] = model, training
        out["BERT"] = bert
        return out
