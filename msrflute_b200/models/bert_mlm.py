"""Masked-language-model fine-tuning with HuggingFace encoders (ref. ``experiments/mlm_bert/model.py``).

Parity: ``model_config.BERT.{model, training}`` arguments (``model_name``, ``model_name_or_path``, ``cache_dir``,
``use_fast_tokenizer``, ``gradient_accumulation_steps``, ``label_smoothing_factor``, ``seed``, ``batch_size`` …), loss =
the HF MLM loss (optionally label-smoothed) divided by ``gradient_accumulation_steps`` (ref :186-245), RoBERTa inputs
stripped of ``attention_mask`` / ``special_tokens_mask`` (ref :218-223), ``inference`` returns eval loss and masked-token
accuracy (ref :247-366 → ``ComputeMetrics``, ``utils/trainer_utils.py:64-86``).

Differences: evaluation is ONE forward per batch (the reference runs the HF ``prediction_loop`` machinery and copies all
logits — B·S·30522 floats — to the host to compute an argmax); accuracy is computed on the device.  Without network
access ``from_pretrained`` cannot download: a local ``model_name_or_path`` directory is used when it exists, otherwise
the architecture named by ``model_name`` is instantiated from its config with random weights (the benchmark setting:
"random-init weights of that architecture").  Attention runs through PyTorch SDPA (flash kernels on sm_100).
"""
import logging
import math
import os

import torch as T

from ..core.model import BaseModel
from ..utils import print_rank

_ARCH = {
    # name fragment -> (hf model_type, config overrides)
    "tiny": ("bert", dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                          vocab_size=1000, max_position_embeddings=128)),
    "roberta-large": ("roberta", dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                      intermediate_size=4096, vocab_size=50265, max_position_embeddings=514,
                                      type_vocab_size=1)),
    "roberta": ("roberta", dict(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1)),
    "bert-large": ("bert", dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)),
    "bert": ("bert", dict()),
}


def build_hf_mlm(model_args):
    from transformers import AutoConfig, AutoModelForMaskedLM
    path = model_args.get("model_name_or_path", model_args["model_name"])
    if isinstance(path, str) and os.path.isdir(path):
        return AutoModelForMaskedLM.from_pretrained(path, local_files_only=True)
    name = str(model_args["model_name"]).lower()
    for frag, (mtype, over) in _ARCH.items():
        if frag in name:
            over = dict(over, **(model_args.get("config_overrides", {}) or {}))
            cfg = AutoConfig.for_model(mtype, **over)
            cfg._attn_implementation = "sdpa"
            if model_args.get("tcgen05_attention", True):
                # hand-written flash-style attention (csrc/attention_tc.cu) through transformers' attention registry;
                # shapes the kernel does not cover (head_dim != 64, S > 512, CPU) fall back to SDPA inside the hook
                try:
                    from ..ops.attention_ops import register_hf
                    cfg._attn_implementation = register_hf()
                    return AutoModelForMaskedLM.from_config(cfg)
                except Exception as exc:         # noqa: BLE001 - an incompatible transformers version keeps SDPA
                    logging.getLogger(__name__).warning("tcgen05 attention hook unavailable (%s); using SDPA", exc)
                    cfg._attn_implementation = "sdpa"
            return AutoModelForMaskedLM.from_config(cfg)
    raise ValueError("unknown architecture {!r}: give a local model_name_or_path".format(model_args["model_name"]))


class LabelSmoother:
    """NLL with uniform label smoothing over the vocabulary, ignoring ``ignore_index`` positions."""

    def __init__(self, epsilon=0.1, ignore_index=-100):
        self.epsilon, self.ignore_index = epsilon, ignore_index

    def __call__(self, model_output, labels):
        logits = model_output["logits"] if isinstance(model_output, dict) else model_output[0]
        logp = -T.nn.functional.log_softmax(logits.float(), dim=-1)
        labels = labels.to(logits.device)
        pad = labels.eq(self.ignore_index)
        nll = logp.gather(-1, labels.clamp(min=0).unsqueeze(-1)).squeeze(-1).masked_fill(pad, 0.0)
        smooth = logp.sum(-1).masked_fill(pad, 0.0)
        n = (~pad).sum().clamp(min=1)
        return (1 - self.epsilon) * nll.sum() / n + self.epsilon * smooth.sum() / (n * logits.shape[-1])


class BERT(BaseModel):
    def __init__(self, model_config, **kwargs):
        super().__init__()
        args = model_config["BERT"]
        model_args, training_args = args["model"], args.get("training", {}) or {}
        T.manual_seed(int(training_args.get("seed", 12345)))
        self.gradient_accumulation_steps = model_args.get("gradient_accumulation_steps", 1)
        self.batch_size = training_args.get("batch_size", 8)
        self.model_name = model_args["model_name"]
        eps = training_args.get("label_smoothing_factor", 0) or 0
        self.label_smoother = LabelSmoother(eps) if eps != 0 else None
        self.model = build_hf_mlm(model_args)
        vocab = model_args.get("vocab_size", None)
        if vocab:
            self.model.resize_token_embeddings(int(vocab))
        if model_args.get("adapter", False) and hasattr(self.model, "add_adapter"):
            self.model.add_adapter("FLUTE")
            self.model.train_adapter("FLUTE")
        # fraction of the tokens of a batch that can carry an MLM label (mlm_probability 0.15 + a wide margin); 1.0 = HF's
        # dense head.  Rows beyond the capacity would be silently ignored, so the margin is ~17 sigma at 4096 tokens.
        self.mlm_head_rows = float(model_args.get("mlm_head_rows", 0.25))
        self.tc_layers = 0
        if model_args.get("tcgen05_linear", True):
            from ..ops.linear_ops import swap_linear_modules
            # encoder GEMMs (QKV / attention-out / FFN) → hand-written tcgen05 kernel; the tied vocabulary
            # decoder keeps its own module (weight tying is implemented by HF on that specific attribute)
            enc = getattr(self.model, "bert", None) or getattr(self.model, "roberta", None)
            if enc is not None:
                self.tc_layers = swap_linear_modules(enc)
        self.fused_layer_norms = 0
        if model_args.get("fused_layer_norm", True):
            from ..ops.norm_ops import swap_layer_norm_modules
            self.fused_layer_norms = swap_layer_norm_modules(self.model)     # one fwd + one bwd kernel per LayerNorm
        n = sum(p.numel() for p in self.model.parameters())
        print_rank("mlm_bert: {} with {:.1f}M parameters".format(self.model_name, n / 1e6), logging.INFO)

    def get_model(self):
        return self.model

    def _prepare_inputs(self, inputs):
        dev = next(self.model.parameters()).device
        out = {k: (v.to(dev, non_blocking=True) if isinstance(v, T.Tensor) else v) for k, v in inputs.items()
               if k in ("input_ids", "attention_mask", "token_type_ids", "labels", "special_tokens_mask", "position_ids")}
        out.pop("special_tokens_mask", None)
        if "roberta" in str(self.model_name).lower():
            out.pop("attention_mask", None)
        return out

    def forward(self, inputs):
        return self.model(**self._prepare_inputs(inputs))

    # ---- masked-rows MLM head ------------------------------------------------------------------------------------------
    # HF computes vocabulary logits for EVERY position and lets CrossEntropyLoss ignore the ~85 % whose label is -100
    # (/root/reference/experiments/mlm_bert/model.py:168-189 calls that model as is).  The loss only needs the masked
    # rows: select them (fixed capacity, so the step stays CUDA-graph capturable), run transform + decoder on those rows
    # only and feed the fused softmax-CE kernel.  The 768 x 30522 projection — the largest GEMM of the step, forward and
    # backward — shrinks by ~5x and the [tokens, vocab] fp32 logits are never materialised.
    def _sparse_head_parts(self):
        m = self.model
        enc = getattr(m, "bert", None) or getattr(m, "roberta", None)
        head = getattr(m, "cls", None) or getattr(m, "lm_head", None)
        return (enc, head) if enc is not None and head is not None else (None, None)

    def _sparse_mlm_loss(self, inputs):
        enc, head = self._sparse_head_parts()
        labels = inputs["labels"]
        enc_in = {k: v for k, v in inputs.items() if k != "labels"}
        seq = enc(**enc_in, return_dict=True)[0]
        n = labels.numel()
        cap = min(n, max(8, int(math.ceil(self.mlm_head_rows * n))))
        flat = labels.reshape(-1)
        valid = flat != -100
        # rows with a label first (stable), a fixed number of them: no host sync, static shapes
        order = T.argsort(valid.to(T.int8), descending=True, stable=True)[:cap]
        rows = seq.reshape(n, seq.shape[-1]).index_select(0, order)
        lab = flat.index_select(0, order)
        logits = head(rows)
        if self.label_smoother is not None:
            return self.label_smoother({"logits": logits}, lab)
        from ..ops import misc_ops
        per_row = misc_ops.softmax_cross_entropy(logits.float(), lab, ignore_index=-100)
        return per_row.sum() / (lab != -100).sum().clamp(min=1)

    def compute_loss(self, inputs, return_outputs=False):
        inputs = self._prepare_inputs(inputs)
        if (not return_outputs and self.training and self.mlm_head_rows < 1.0 and "labels" in inputs
                and self._sparse_head_parts()[0] is not None):
            return self._sparse_mlm_loss(inputs)
        labels = inputs["labels"] if self.label_smoother is not None and "labels" in inputs else None
        outputs = self.model(**inputs)
        loss = self.label_smoother(outputs, labels) if labels is not None else outputs["loss"]
        return (loss, outputs) if return_outputs else loss

    def loss(self, inputs):
        return self.compute_loss(inputs) / self.gradient_accumulation_steps

    @staticmethod
    def masked_accuracy(logits, labels):
        mask = labels != -100
        hit = (logits.argmax(-1) == labels) & mask
        return hit.sum().float() / mask.sum().clamp(min=1)

    def loss_and_metrics(self, inputs):
        self.model.eval()
        with T.no_grad():
            loss, outputs = self.compute_loss(inputs, return_outputs=True)
            labels = inputs["labels"].to(outputs["logits"].device)
            acc = self.masked_accuracy(outputs["logits"], labels)
        return loss.mean().detach(), {"output": None, "acc": acc, "batch_size": int(labels.shape[0])}

    def inference(self, inputs, ignore_keys=None, metric_key_prefix="eval"):
        loss, m = self.loss_and_metrics(inputs)
        return {"output": loss.item(), "acc": m["acc"].item(), "batch_size": m["batch_size"]}
