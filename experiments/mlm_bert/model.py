"""mlm_bert task model (HuggingFace encoder + MLM head)."""
from msrflute_b200.models.bert_mlm import BERT, LabelSmoother, build_hf_mlm  # noqa: F401
