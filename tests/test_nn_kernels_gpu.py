"""Embedding, dropout and BatchNorm kernels (csrc/nn_kernels.cu) against the PyTorch fp32 ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("V,D,shape,pad", [(90, 8, (4, 80), 0), (10000, 160, (7, 25), None), (30522, 768, (2, 16), 0)])
def test_embedding_fwd_bwd(V, D, shape, pad):
    from msrflute_b200.ops import nn_ops
    torch.manual_seed(0)
    w = torch.randn(V, D, device="cuda", requires_grad=True)
    idx = torch.randint(0, V, shape, device="cuda")
    idx[0, :3] = 0
    dy = torch.randn(*shape, D, device="cuda")
    out = nn_ops.embedding(idx, w, pad)
    out.backward(dy)
    g1 = w.grad.clone()
    w.grad = None
    ref = F.embedding(idx, w, padding_idx=pad)
    ref.backward(dy)
    assert torch.equal(out, ref)
    assert torch.allclose(g1, w.grad, atol=1e-5, rtol=1e-5)


def test_dropout_statistics_and_backward_mask():
    from msrflute_b200.ops import nn_ops
    torch.manual_seed(0)
    d = nn_ops.Dropout(0.25).cuda().train()
    x = torch.ones(1 << 20, device="cuda", requires_grad=True)
    y = d(x)
    keep = (y != 0)
    assert abs(float(keep.float().mean()) - 0.75) < 5e-3
    assert torch.allclose(y[keep], torch.full_like(y[keep], 1 / 0.75))
    y.sum().backward()
    assert torch.equal(x.grad != 0, keep) and torch.allclose(x.grad[keep], torch.full_like(y[keep], 1 / 0.75))
    y2 = d(x)
    assert not torch.equal(y2 != 0, keep)                     # the device counter moved on
    # a captured graph draws a new mask on every replay
    xs = torch.ones(4096, device="cuda")
    out = torch.empty_like(xs)
    d(xs)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out.copy_(d(xs))
    g.replay(); torch.cuda.synchronize(); m1 = (out != 0).clone()
    g.replay(); torch.cuda.synchronize(); m2 = (out != 0).clone()
    assert not torch.equal(m1, m2)
    d.eval()
    assert torch.equal(d(xs), xs)


def test_dropout_under_vmap_grad():
    """The engine's wave step is vmap(grad_and_value(loss)): the dropout op must carry its vmap rule."""
    from torch.func import functional_call, grad_and_value, vmap
    from msrflute_b200.ops import nn_ops
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), nn_ops.Dropout(0.5), torch.nn.Linear(32, 4)).cuda().train()
    params = {k: v.detach().unsqueeze(0).repeat(3, *([1] * v.dim())) for k, v in m.named_parameters()}
    x = torch.randn(3, 5, 16, device="cuda")

    def loss(p, xb):
        return functional_call(m, p, (xb,)).pow(2).mean()

    grads, vals = vmap(grad_and_value(loss), in_dims=(0, 0), randomness="different")(params, x)
    assert all(torch.isfinite(g).all() for g in grads.values()) and torch.isfinite(vals).all()


@pytest.mark.parametrize("N,C,H,res,relu", [(20, 64, 8, False, True), (20, 128, 4, True, True), (3, 512, 1, True, False),
                                            (7, 16, 5, False, False)])
def test_batch_norm_train_matches_torch(N, C, H, res, relu):
    from msrflute_b200.ops import nn_ops
    torch.manual_seed(1)
    x = torch.randn(N, C, H, H, device="cuda", requires_grad=True)
    r = torch.randn(N, C, H, H, device="cuda", requires_grad=True) if res else None
    g = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    rm1, rv1 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm2, rv2 = rm1.clone(), rv1.clone()
    dy = torch.randn(N, C, H, H, device="cuda")
    y = nn_ops.batch_norm_train(x, g, b, rm1, rv1, 0.1, 1e-5, residual=r, relu=relu)
    y.backward(dy)
    got = [y.detach(), x.grad.clone(), g.grad.clone(), b.grad.clone()] + ([r.grad.clone()] if res else [])
    for t in [x, g, b] + ([r] if res else []):
        t.grad = None
    yr = F.batch_norm(x, rm2, rv2, g, b, True, 0.1, 1e-5)
    if res:
        yr = yr + r
    if relu:
        yr = F.relu(yr)
    yr.backward(dy)
    want = [yr.detach(), x.grad, g.grad, b.grad] + ([r.grad] if res else [])
    for a, e in zip(got, want):
        assert torch.allclose(a, e, atol=2e-4, rtol=2e-4), float((a - e).abs().max())
    assert torch.allclose(rm1, rm2, atol=1e-5) and torch.allclose(rv1, rv2, atol=1e-4, rtol=1e-4)


def test_gru_layer_runs_its_gemms_on_the_tcgen05_kernel():
    """GEMM-shaped GRU batches (the nlg_gru task trains with 2048 rows): both projections go through ops.linear_ops
    (bf16 operands) — same hidden states / gradients as the fp32 module within bf16 tolerance."""
    from msrflute_b200.models.gru_lm import GRU2
    from msrflute_b200.ops import _ext
    torch.manual_seed(0)
    g = GRU2(160, 512).cuda()
    x = torch.randn(512, 6, 160, device="cuda") * 0.5
    n0 = _ext.LAUNCH_COUNTER["n"]
    hs, h = g(x)
    assert _ext.LAUNCH_COUNTER["n"] - n0 >= 7            # 1 input GEMM + 6 recurrent GEMMs (+ gate kernels)
    loss = (hs * hs).mean()
    loss.backward()
    g_tc = g.w_hh.weight.grad.clone()
    g.zero_grad()
    g.tcgen05_linear = False
    hs_ref, _ = g(x)
    (hs_ref * hs_ref).mean().backward()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))
    assert rel(hs, hs_ref) < 2e-2 and rel(g_tc, g.w_hh.weight.grad) < 5e-2, (rel(hs, hs_ref), rel(g_tc, g.w_hh.weight.grad))
