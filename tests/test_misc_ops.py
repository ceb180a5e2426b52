"""CPU reference paths of ops.misc_ops / quant_ops (the oracle the CUDA kernels are tested against on GPU)."""
import torch

from msrflute_b200.ops import misc_ops, quant_ops
from msrflute_b200.utils import compute_grad_cosines


def test_local_dp_reference_clip_and_normalise():
    torch.manual_seed(0)
    g = torch.randn(10000) * 3
    h = g.clone()
    n = misc_ops.local_dp_(h, 2.0, 0.0, True)
    assert abs(n.item() - g.norm().item()) < 1e-4 and abs(h.norm().item() - 2.0) < 1e-4
    h = g.clone()
    misc_ops.local_dp_(h, 2.0, 0.5, False, seed=3)
    noise = h - g * (2.0 / g.norm())
    assert abs(noise.std().item() - 0.5) < 0.02


def test_softmax_cross_entropy_fallback_and_cosine():
    x = torch.randn(5, 7, requires_grad=True)
    t = torch.tensor([1, 2, 3, 4, 5])
    loss = misc_ops.softmax_cross_entropy(x, t)
    assert torch.allclose(loss, torch.nn.functional.cross_entropy(x, t, reduction="none"))
    a, b = torch.randn(100), torch.randn(100)
    s = misc_ops.cosine_stats(a, b)
    assert torch.allclose(s, torch.stack([a @ b, a @ a, b @ b]))
    assert abs(misc_ops.cosine(a, -a).item() + 1.0) < 1e-6
    cs = compute_grad_cosines([[a], [b]], [a])
    assert abs(cs[0] - 1.0) < 1e-6 and abs(cs[1] - float(a @ b / (a.norm() * b.norm()))) < 1e-6


def test_quantize_segments_reference_levels_and_sparsity():
    torch.manual_seed(1)
    flat = torch.randn(5000)
    out = quant_ops.quantize_segments_(flat.clone(), [(0, 3000), (3000, 2000)], 4, 0.5)
    for o, n in ((0, 3000), (3000, 2000)):
        seg = out[o:o + n]
        assert abs((seg == 0).float().mean().item() - 0.5) < 0.01
        assert seg[seg != 0].unique().numel() <= 16


def test_attention_ops_cpu_fallback_and_mask_conversion():
    """ops.attention_ops off the GPU: SDPA fallback with the [B, S, H, D] output contract, key-bias handling, and the
    compact padding-mask builder registered with transformers."""
    import torch
    from msrflute_b200.ops import attention_ops as A
    torch.manual_seed(0)
    q, k, v = (torch.randn(2, 3, 10, 16) for _ in range(3))
    bias = torch.zeros(2, 10)
    bias[1, 7:] = float("-inf")
    out = A.attention(q, k, v, key_bias=bias)
    ref = A.attention_reference(q, k, v, key_bias=bias)
    assert out.shape == (2, 10, 3, 16) and torch.allclose(out, ref, atol=1e-5)
    # HF-style masks: a compact [B, 1, 1, S] bool mask becomes a key bias; a materialised [B, 1, S, S] one does not
    compact = torch.ones(2, 1, 1, 10, dtype=torch.bool)
    compact[1, :, :, 7:] = False
    kb, ok = A._key_bias_from_mask(compact, q)
    assert ok and torch.equal(torch.isinf(kb), ~compact[:, 0, 0, :])
    assert A._key_bias_from_mask(compact.expand(2, 1, 10, 10).contiguous(), q) == (None, False)
    assert A._key_bias_from_mask(None, q) == (None, True)
    o2, w = A.hf_attention_forward(None, q, k, v, compact, dropout=0.0, scaling=0.25)
    assert w is None and torch.allclose(o2, A.attention_reference(q, k, v, key_bias=bias, scale=0.25), atol=1e-5)
    from transformers import masking_utils as mu
    am = torch.ones(2, 10, dtype=torch.long)
    am[1, 7:] = 0
    m = A._padding_only_mask(2, 10, 10, mask_function=mu.bidirectional_mask_function, attention_mask=am)
    assert m.shape == (2, 1, 1, 10) and m.dtype == torch.bool and torch.equal(m[:, 0, 0, :], am.bool())
    assert A._padding_only_mask(2, 10, 10, mask_function=mu.bidirectional_mask_function, attention_mask=None) is None
