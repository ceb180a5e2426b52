"""Recurrent-cell pointwise math fused into single kernels.

``gru_cell(gi, gh, h)``   r = σ(gi_r+gh_r); z = σ(gi_z+gh_z); n = tanh(gi_n + r·gh_n); h' = n + z·(h − n)
                          (the reference's hand-rolled GRU, ``experiments/nlg_gru/model.py:19-30``)
``lstm_cell(gates, c)``   i,f,g,o = chunk(gates); c' = σ(f)·c + σ(i)·tanh(g); h' = σ(o)·tanh(c')
                          (``nn.LSTM`` gate order; ``experiments/nlp_rnn_fedshakespeare/model.py:18-23``)

CUDA: ``csrc/rnn_kernels.cu`` (one launch forward, one backward; gate pre-activations are recomputed in the backward
from the saved inputs instead of storing 3–4 activation tensors).  PyTorch reference below is the CPU path / oracle.
"""
import torch

from . import _ext


def _gru_cell_ref(gi, gh, h):
    H = h.shape[-1]
    i_r, i_z, i_n = gi.split(H, dim=-1)
    h_r, h_z, h_n = gh.split(H, dim=-1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return n + z * (h - n)


def _lstm_cell_ref(gates, c):
    H = c.shape[-1]
    i, f, g, o = gates.split(H, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c2), c2


class _GRUCellFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, gh, h):
        gi, gh, h = gi.contiguous(), gh.contiguous(), h.contiguous()
        out = _ext.load().gru_cell_fwd(gi, gh, h)
        _ext.count_launch(1)
        ctx.save_for_backward(gi, gh, h)
        return out

    @staticmethod
    def backward(ctx, dh):
        gi, gh, h = ctx.saved_tensors
        dgi, dgh, dhp = _ext.load().gru_cell_bwd(dh.contiguous(), gi, gh, h)
        _ext.count_launch(1)
        return dgi, dgh, dhp


class _LSTMCellFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, c):
        gates, c = gates.contiguous(), c.contiguous()
        h2, c2 = _ext.load().lstm_cell_fwd(gates, c)
        _ext.count_launch(1)
        ctx.save_for_backward(gates, c, c2)
        return h2, c2

    @staticmethod
    def backward(ctx, dh, dc):
        gates, c, c2 = ctx.saved_tensors
        dgates, dcp = _ext.load().lstm_cell_bwd(dh.contiguous(), dc.contiguous(), gates, c, c2)
        _ext.count_launch(1)
        return dgates, dcp


def _has(name):
    ext = _ext.load() if torch.cuda.is_available() else None
    return ext is not None and hasattr(ext, name)


def gru_cell(gi, gh, h):
    if gi.is_cuda and gi.dtype == torch.float32 and _has("gru_cell_fwd") and not torch._C._functorch.is_batchedtensor(gi):
        return _GRUCellFn.apply(gi, gh, h)
    return _gru_cell_ref(gi, gh, h)


def lstm_cell(gates, c):
    if gates.is_cuda and gates.dtype == torch.float32 and _has("lstm_cell_fwd") \
            and not torch._C._functorch.is_batchedtensor(gates):
        return _LSTMCellFn.apply(gates, c)
    return _lstm_cell_ref(gates, c)
