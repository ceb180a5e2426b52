"""Fused arena AdamW client step (clip + statistics + AdamW + zero-grad) against the framework's AdamW optimizer."""
import pytest
import torch

from msrflute_b200.ops import arena_ops
from msrflute_b200.utils.optimizers import AdamW


def _run(device, steps=4, clip=0.0):
    torch.manual_seed(0)
    P = 1024
    w0 = torch.randn(P)
    grads = [torch.randn(P) * (0.5 + i) for i in range(steps)]
    # optimizer path
    p = torch.nn.Parameter(w0.clone())
    opt = AdamW([p], lr=1e-2, weight_decay=0.05, eps=1e-6)
    for g in grads:
        p.grad = g.clone()
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([p], clip)
        opt.step()
    # fused arena path
    w = w0.clone().to(device).view(1, P)
    g = torch.zeros(1, P, device=device)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    step = torch.zeros(1, dtype=torch.int32, device=device)
    hyper = arena_ops.make_hyper(1, device)
    hyper.copy_(torch.tensor([[1e-2, clip, 0.05, 0.0]]))
    stats = torch.zeros(1, 4, device=device)
    for gi in grads:
        g.copy_(gi.view(1, P))
        step.add_(1)
        arena_ops.fused_client_adamw(w, g, m, v, step, hyper, stats, n_logical=P, betas=(0.9, 0.999), eps=1e-6)
        assert float(g.abs().max()) == 0.0
    return p.detach(), w.view(-1).cpu(), stats.cpu()


@pytest.mark.parametrize("clip", [0.0, 1.0])
def test_fused_adamw_reference_path_matches_optimizer(clip):
    want, got, stats = _run("cpu", clip=clip)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert float(stats[0, arena_ops.S_COUNT]) == 4 * 1024


@pytest.mark.gpu
@pytest.mark.parametrize("clip", [0.0, 1.0])
def test_fused_adamw_kernel_matches_optimizer(clip):
    want, got, stats = _run("cuda", clip=clip)
    assert torch.allclose(got, want, rtol=2e-5, atol=2e-6), float((got - want).abs().max())
