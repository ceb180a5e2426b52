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
