"""Persistent LSTM layer: one launch per layer and direction of time, W_hh resident in shared memory across
all time steps, 8-CTA clusters exchanging h through distributed shared memory (``csrc/lstm_kernels.cu``).

``lstm_layer(x, w_ih, w_hh, b_ih, b_hh, h0, c0)`` has ``torch.nn.LSTM``'s single-layer, batch-first semantics (gate order
i, f, g, o).  The input projection ``x·W_ihᵀ + b`` of ALL time steps is one GEMM; the backward's ``dW_hh``, ``dW_ih`` and
``dx`` are GEMMs over the saved ``[B·T, ·]`` tensors.  CPU / unsupported hidden sizes use the plain PyTorch loop below,
which is also the test oracle."""
import torch

from . import _ext


def lstm_layer_reference(x, w_ih, w_hh, b_ih, b_hh, h0=None, c0=None):
    """Plain PyTorch time loop (fp32): returns ``(hs [B,T,H], (hT, cT))``."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = h0 if h0 is not None else x.new_zeros(B, H)
    c = c0 if c0 is not None else x.new_zeros(B, H)
    gx = torch.matmul(x, w_ih.t())
    if b_ih is not None:
        gx = gx + b_ih + b_hh
    outs = []
    for t in range(T):
        g = gx[:, t] + torch.matmul(h, w_hh.t())
        i, f, gg, o = g.split(H, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1), (h, c)


def _supported(x, w_hh):
    if not (x.is_cuda and x.dtype == torch.float32 and w_hh.dtype == torch.float32):
        return False
    ext = _ext.load()
    return ext is not None and hasattr(ext, "lstm_layer_fwd") and bool(ext.lstm_supported(int(w_hh.shape[1])))


class _LSTMLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, h0, c0):
        ext = _ext.load()
        B, T, _ = x.shape
        H = w_hh.shape[1]
        gx = torch.matmul(x.reshape(B * T, -1), w_ih.t())
        if b_ih is not None:
            gx = gx + (b_ih + b_hh)
        gx = gx.view(B, T, 4 * H).contiguous()
        whh = w_hh.contiguous()
        hs, gates, cs = ext.lstm_layer_fwd(gx, whh, h0, c0)
        _ext.count_launch(1)
        ctx.save_for_backward(x, w_ih, whh, hs, gates, cs, h0, c0)
        ctx.has_bias = b_ih is not None
        return hs, hs[:, -1].contiguous(), cs[:, -1].contiguous()

    @staticmethod
    def backward(ctx, dhs, dhT, dcT):
        x, w_ih, whh, hs, gates, cs, h0, c0 = ctx.saved_tensors
        ext = _ext.load()
        B, T, H = hs.shape
        dhs = dhs.contiguous() if dhs is not None else torch.zeros_like(hs)
        dgx, dh0, dc0 = ext.lstm_layer_bwd(dhs, gates, cs, whh, c0, dhT, dcT)
        _ext.count_launch(1)
        d2 = dgx.view(B * T, 4 * H)
        # h_{t-1} for every step: [h0, hs[:, :-1]]
        hprev = torch.cat([(h0 if h0 is not None else hs.new_zeros(B, H)).unsqueeze(1), hs[:, :-1]], dim=1)
        dw_hh = torch.matmul(d2.t(), hprev.reshape(B * T, H))
        dw_ih = torch.matmul(d2.t(), x.reshape(B * T, -1))
        dx = torch.matmul(d2, w_ih).view(x.shape) if ctx.needs_input_grad[0] else None
        db = d2.sum(dim=0) if ctx.has_bias else None
        return (dx, dw_ih, dw_hh, db, db, dh0 if h0 is not None else None, dc0 if c0 is not None else None)


def lstm_layer(x, w_ih, w_hh, b_ih=None, b_hh=None, h0=None, c0=None):
    """``x [B, T, I]`` → ``(hs [B, T, H], (h_T, c_T))``."""
    if _supported(x, w_hh):
        hs, hT, cT = _LSTMLayerFn.apply(x.contiguous(), w_ih, w_hh, b_ih, b_hh, h0, c0)
        return hs, (hT, cT)
    return lstm_layer_reference(x, w_ih, w_hh, b_ih, b_hh, h0, c0)


class LSTM(torch.nn.Module):
    """Drop-in for ``torch.nn.LSTM(input_size, hidden_size, num_layers, batch_first=True)`` (unidirectional) with the
    same parameter names (``weight_ih_l0`` …, so checkpoints are interchangeable) running on the persistent kernel.
    Parameters are ordinary ``nn.Parameter``s — no cuDNN weight flattening, so they can live in a flat arena."""

    def __init__(self, input_size, hidden_size, num_layers=1, batch_first=True, bias=True):
        super().__init__()
        assert batch_first, "batch_first layout only"
        self.input_size, self.hidden_size, self.num_layers, self.bias = input_size, hidden_size, num_layers, bias
        k = 1.0 / hidden_size ** 0.5
        for layer in range(num_layers):
            i = input_size if layer == 0 else hidden_size
            self.register_parameter("weight_ih_l{}".format(layer), torch.nn.Parameter(torch.empty(4 * hidden_size, i).uniform_(-k, k)))
            self.register_parameter("weight_hh_l{}".format(layer), torch.nn.Parameter(torch.empty(4 * hidden_size, hidden_size).uniform_(-k, k)))
            if bias:
                self.register_parameter("bias_ih_l{}".format(layer), torch.nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k)))
                self.register_parameter("bias_hh_l{}".format(layer), torch.nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k)))

    def forward(self, x, hx=None):
        hT, cT = [], []
        out = x
        for layer in range(self.num_layers):
            h0 = hx[0][layer] if hx is not None else None
            c0 = hx[1][layer] if hx is not None else None
            out, (h, c) = lstm_layer(out, getattr(self, "weight_ih_l{}".format(layer)),
                                     getattr(self, "weight_hh_l{}".format(layer)),
                                     getattr(self, "bias_ih_l{}".format(layer)) if self.bias else None,
                                     getattr(self, "bias_hh_l{}".format(layer)) if self.bias else None, h0, c0)
            hT.append(h)
            cT.append(c)
        return out, (torch.stack(hT), torch.stack(cT))
