"""Next-word-prediction GRU (Reddit) — ref. ``experiments/nlg_gru/model.py``.

Architecture parity: tied embedding ``table`` used to embed and (transposed, + ``unembedding_bias``) to un-embed
(ref :39-54), a single-layer GRU with separate input / hidden projections ``w_ih`` / ``w_hh`` (ref :11-36,
``h' = n + z·(h − n)``), a bias-free ``squeeze`` Linear hidden→embed (ref :68), loss over all positions *including*
the prediction from the zero initial state (the hidden sequence starts with h0, ref :31-36,92-98), pad id −1,
OOV id 0 (predictions of id 0 never count as correct unless ``OOV_correct``).

B200-first: the reference's python time loop issues ~8 small kernels per step.  Here the input projection for
ALL time steps is one GEMM up front (``X·W_ihᵀ``) and the recurrent part per step is one GEMM + one fused gate kernel
(``ops.rnn_ops.gru_cell``); with GEMM-shaped batches (≥ 256 rows — the task's 2048) both GEMMs, forward and backward,
run on the hand-written tcgen05 kernel (``ops.linear_ops``, bf16 operands / fp32 accumulate) instead of cuBLAS.
"""
from typing import Tuple

import torch as T
from torch import Tensor

from ..core.model import BaseModel
from ..ops import rnn_ops
from ..utils import softmax


class GRU2(T.nn.Module):
    def __init__(self, input_size, hidden_size, input_bias=True, hidden_bias=True):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.w_ih = T.nn.Linear(input_size, 3 * hidden_size, input_bias)
        self.w_hh = T.nn.Linear(hidden_size, 3 * hidden_size, hidden_bias)

    #: run the two projections on the hand-written tcgen05 GEMM (bf16 operands, fp32 accumulate) when the problem is
    #: GEMM-shaped: the reference task trains with 2048-row batches, i.e. [2048 x 512] x [512 x 1536] per time step
    TC_MIN_ROWS = 256

    def _proj(self, lin, x):
        if x.is_cuda and x.shape[0] >= self.TC_MIN_ROWS and getattr(self, "tcgen05_linear", True):
            from ..ops import linear_ops
            if linear_ops.tc_available(x, lin.weight):
                return linear_ops.linear(x, lin.weight, lin.bias).to(x.dtype)
        return lin(x)

    def forward(self, input: Tensor) -> Tuple[Tensor, Tensor]:
        B, L, _ = input.shape
        gi_all = self._proj(self.w_ih, input.reshape(B * L, -1)).view(B, L, -1)   # one GEMM for every time step
        h = input.new_zeros(B, self.hidden_size)
        hs = [h]
        for t in range(L):
            h = rnn_ops.gru_cell(gi_all[:, t], self._proj(self.w_hh, h), h)        # fused r/z/n gates + state update
            hs.append(h)
        return T.stack(hs, dim=1), h


class Embedding(T.nn.Module):
    def __init__(self, vocab_size, embedding_size):
        super().__init__()
        self.vocab_size, self.embedding_size = vocab_size, embedding_size
        self.table = T.nn.Parameter(T.zeros((vocab_size, embedding_size)))
        self.unembedding_bias = T.nn.Parameter(T.zeros(vocab_size))
        delta = (3 / embedding_size) ** 0.5
        T.nn.init.uniform_(self.table, -delta, delta)

    def forward(self, input: Tensor, embed: bool) -> Tensor:
        if embed:
            return T.nn.functional.embedding(input, self.table)
        return input @ self.table.t() + self.unembedding_bias


class GRU(BaseModel):
    def __init__(self, model_config, OOV_correct=False, dropout=0.0, topK_results=1, wantLogits=False, **kwargs):
        super().__init__()
        self.vocab_size = model_config["vocab_size"]
        self.embedding_size = model_config["embed_dim"]
        self.hidden_size = model_config["hidden_dim"]
        self.embedding = Embedding(self.vocab_size, self.embedding_size)
        self.rnn = GRU2(self.embedding_size, self.hidden_size, True, True)
        self.squeeze = T.nn.Linear(self.hidden_size, self.embedding_size, bias=False)
        self.OOV_correct = model_config.get("OOV_correct", OOV_correct)
        self.topK_results = model_config.get("topK_results", topK_results)
        self.dropout = dropout
        self.wantLogits = model_config.get("wantLogits", wantLogits)
        self.drop_layer = T.nn.Dropout(p=dropout) if dropout > 0.0 else None

    def _tokens(self, input):
        x = input["x"] if isinstance(input, dict) else input
        return x.to(self.embedding.table.device)

    def forward(self, input) -> Tuple[Tensor, Tensor]:
        x = self._tokens(input)
        hiddens, state = self.rnn(self.embedding(x, True))
        if self.drop_layer is not None:
            hiddens = self.drop_layer(hiddens)
        return self.embedding(self.squeeze(hiddens), False), state

    def _preds_targets(self, input):
        x = self._tokens(input)
        mask = x >= 0
        x = x * mask.long()
        output, _ = self.forward(x[:, :-1])
        m = mask.reshape(-1)
        return output.reshape(-1, self.vocab_size)[m], x.reshape(-1)[m], x.shape[0]

    def token_logits(self, input):
        """Per-position logits ``[b, s, v]`` (position t predicts token t from tokens < t: the recurrence starts from
        a zero step), their targets and the non-pad mask — what the privacy leakage metric needs."""
        x = self._tokens(input)
        mask = x >= 0
        x = x * mask.long()
        output, _ = self.forward(x[:, :-1])
        return output, x, mask

    def loss(self, input) -> Tensor:
        preds, targets, _ = self._preds_targets(input)
        return T.nn.functional.cross_entropy(preds, targets)

    def _metrics(self, preds, targets, bsz, want_output):
        probs_topK, preds_topK = T.topk(preds, self.topK_results, sorted=True, dim=1)
        top1 = preds_topK[:, 0]
        hit = top1.eq(targets)
        if not self.OOV_correct:
            hit = hit & (top1 != 0)               # an OOV prediction never counts, even when it matches
        output = None
        if want_output and self.wantLogits:
            output = {"probabilities": softmax(probs_topK.detach().cpu().numpy(), axis=1),
                      "predictions": preds_topK.detach().cpu().numpy(), "labels": targets.detach().cpu().numpy()}
        return {"output": output, "acc": hit.float().mean(), "batch_size": bsz}

    def inference(self, input):
        preds, targets, bsz = self._preds_targets(input)
        out = self._metrics(preds, targets, bsz, True)
        out["acc"] = out["acc"].item()
        return out

    def loss_and_metrics(self, input):
        preds, targets, bsz = self._preds_targets(input)
        return T.nn.functional.cross_entropy(preds, targets), self._metrics(preds, targets, bsz, True)
