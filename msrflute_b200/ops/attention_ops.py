"""Multi-head attention on the hand-written tcgen05 kernels (``csrc/attention_tc.cu``).

``attention(q, k, v, ...)`` takes ``[B, H, S, D]`` tensors (any strides with a contiguous head dimension — HF's
``view(B, S, H, D).transpose(1, 2)`` is consumed in place through TMA tensor maps) and returns ``[B, S, H, D]``, the layout
the output projection wants.  CUDA path: ``D == 64``, ``S <= 512``, bf16 operands (fp32 inputs are cast once), optional
additive key bias ``[B, S]``, dropout with an in-kernel Philox stream whose seed lives in a device counter (so a captured
CUDA graph draws a fresh mask on every replay and the backward kernel regenerates the forward's mask bit for bit).
Everything else falls back to ``torch.nn.functional.scaled_dot_product_attention``; :func:`attention_reference` is the
plain fp32 oracle used by the tests.  Replaces the SDPA call behind ``/root/reference/experiments/mlm_bert/model.py:
119-125`` (HF ``BertSelfAttention``).
"""
from __future__ import annotations

import math

import torch

from . import _ext

_SEEDS = {}


def _seed_counter(device) -> torch.Tensor:
    """Per-device int64 counter; every attention call with dropout consumes one value (device-side increment)."""
    key = (device.type, device.index)
    t = _SEEDS.get(key)
    if t is None:
        t = torch.tensor([int(torch.empty((), dtype=torch.int64).random_().item()) & ((1 << 62) - 1)], dtype=torch.int64,
                         device=device)
        _SEEDS[key] = t
    return t


def reseed(seed: int, device=None) -> None:
    dev = torch.device(device if device is not None else "cuda")
    if dev.index is None and dev.type == "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    _seed_counter(dev).fill_(int(seed))


def attention_reference(q, k, v, key_bias=None, scale=None, keep_mask=None, p_drop=0.0):
    """fp32 math on ``[B, H, S, D]`` operands -> ``[B, S, H, D]``.  ``keep_mask``: ``[B, H, S, S]`` (1 = keep)."""
    q32, k32, v32 = q.float(), k.float(), v.float()
    scale = (1.0 / math.sqrt(q.shape[-1])) if scale is None else scale
    s = torch.matmul(q32, k32.transpose(-1, -2)) * scale
    if key_bias is not None:
        s = s + key_bias.float()[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    if keep_mask is not None and p_drop > 0:
        p = p * keep_mask.to(p.dtype) / (1.0 - p_drop)
    return torch.matmul(p, v32).transpose(1, 2)


def supported(q: torch.Tensor) -> bool:
    return (_ext.use_cuda_kernels(q) and q.dim() == 4 and q.shape[-1] == 64 and 1 <= q.shape[-2] <= 512
            and q.dtype in (torch.bfloat16, torch.float32, torch.float16))


def _as_operand(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    if t.stride(-1) != 1 or any((s * 2) % 16 for s in t.stride()[:-1]) or t.data_ptr() % 16:
        t = t.contiguous()
    return t


class _AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_bias, scale, p_drop, seed):
        C = _ext.load(required=True)
        qb, kb, vb = _as_operand(q), _as_operand(k), _as_operand(v)
        out, lse = C.attention_fwd(qb, kb, vb, key_bias, float(scale), float(p_drop), seed)
        _ext.count_launch(1)
        ctx.save_for_backward(qb, kb, vb, out, lse, key_bias, seed)
        ctx.scale, ctx.p_drop, ctx.in_dtypes = float(scale), float(p_drop), (q.dtype, k.dtype, v.dtype)
        return out if q.dtype == torch.bfloat16 else out.to(q.dtype)

    @staticmethod
    def backward(ctx, d_out):
        C = _ext.load(required=True)
        qb, kb, vb, out, lse, key_bias, seed = ctx.saved_tensors
        d_o = d_out.to(torch.bfloat16).contiguous()
        dq, dk, dv = C.attention_bwd(qb, kb, vb, out, lse, d_o, key_bias, ctx.scale, ctx.p_drop, seed)
        _ext.count_launch(1)
        tq, tk, tv = ctx.in_dtypes
        # [B, S, H, D] buffers -> [B, H, S, D] views (what q / k / v were)
        return (dq.to(tq).transpose(1, 2), dk.to(tk).transpose(1, 2), dv.to(tv).transpose(1, 2), None, None, None, None)


def attention(q, k, v, key_bias=None, dropout_p: float = 0.0, scale=None, training: bool = True) -> torch.Tensor:
    """``softmax(q k^T * scale + key_bias) v`` with dropout on the probabilities; returns ``[B, S, H, D]``."""
    scale = (1.0 / math.sqrt(q.shape[-1])) if scale is None else float(scale)
    p = float(dropout_p) if training else 0.0
    if supported(q) and k.shape == q.shape and v.shape == q.shape:
        seed = _seed_counter(q.device)
        if p > 0.0:
            used = seed.clone()                  # the value this call's forward AND backward read
            seed.add_(1)
        else:
            used = seed
        if key_bias is not None:
            key_bias = key_bias.to(torch.float32).contiguous()
        return _AttentionFn.apply(q, k, v, key_bias, scale, p, used)
    mask = None if key_bias is None else key_bias[:, None, None, :].to(q.dtype)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p, scale=scale)
    return o.transpose(1, 2)


def dropout_keep_mask(B, H, S, p_drop, seed: torch.Tensor) -> torch.Tensor:
    """The ``[B, H, S, S]`` keep mask the kernels generate for ``seed`` (test oracle)."""
    return _ext.load(required=True).attention_dropout_mask(int(B), int(H), int(S), float(p_drop), seed)


# ---------------------------------------------------------------------------------------------- HF attention interface
HF_NAME = "flute_tcgen05"


def _key_bias_from_mask(attention_mask, q):
    """HF hands the encoder a 4-D mask; a padding mask is constant over heads and query rows -> ``[B, S]`` additive bias.
    Returns (bias or None, ok)."""
    if attention_mask is None:
        return None, True
    m = attention_mask
    if m.dim() != 4 or m.shape[1] != 1 or m.shape[-1] != q.shape[-2]:
        return None, False
    if not (m.shape[2] == 1 or m.stride(2) == 0):
        return None, False
    row = m[:, 0, 0, :]
    if row.dtype == torch.bool:
        return torch.zeros(row.shape, dtype=torch.float32, device=row.device).masked_fill_(~row, float("-inf")), True
    return row.to(torch.float32), True


def hf_attention_forward(module, query, key, value, attention_mask, dropout: float = 0.0, scaling=None, **kwargs):
    """``transformers.AttentionInterface`` entry: (attn_output [B, S, H, D], None)."""
    bias, ok = _key_bias_from_mask(attention_mask, query)
    if ok and supported(query) and key.shape == query.shape:
        return attention(query, key, value, key_bias=bias, dropout_p=dropout, scale=scaling, training=True), None
    o = torch.nn.functional.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask, dropout_p=dropout,
                                                         scale=scaling)
    return o.transpose(1, 2).contiguous(), None


def _padding_only_mask(batch_size, q_length, kv_length, q_offset=0, kv_offset=0, mask_function=None, attention_mask=None,
                       **kwargs):
    """Mask builder registered next to the attention function: an encoder's padding mask stays COMPACT (``[B, 1, 1, S]``
    bool, what the kernel turns into a per-key bias) instead of being materialised as ``[B, 1, S, S]``; anything that is
    not a plain bidirectional padding mask goes through transformers' SDPA mask builder unchanged."""
    from transformers import masking_utils as mu
    if mask_function is mu.bidirectional_mask_function and kv_offset == 0 and q_offset == 0:
        if attention_mask is None:
            return None
        if attention_mask.dim() == 2 and attention_mask.shape[-1] == kv_length:
            return attention_mask[:, None, None, :].to(torch.bool)
    return mu.sdpa_mask(batch_size, q_length, kv_length, q_offset=q_offset, kv_offset=kv_offset, mask_function=mask_function,
                        attention_mask=attention_mask, **kwargs)


def register_hf() -> str:
    """Register the kernel with transformers' attention registry; returns the implementation name to put in the config."""
    from transformers import AttentionInterface
    AttentionInterface.register(HF_NAME, hf_attention_forward)
    try:
        from transformers.masking_utils import AttentionMaskInterface
        AttentionMaskInterface.register(HF_NAME, _padding_only_mask)
    except Exception:                            # noqa: BLE001 - other transformers versions: their default builder is used
        pass
    return HF_NAME
