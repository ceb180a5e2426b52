import math

import pytest
import torch

from msrflute_b200.ops import arena_ops
from msrflute_b200.parallel.arena import ArenaLayout, adopt_module, module_arena


def _rows(S=3, P=4096, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(S, P, generator=g).to(device), torch.randn(S, P, generator=g).to(device) * 3)


def _manual_step(w, g, lr, max_norm, wd):
    norm = g.norm(dim=1)
    coef = torch.where(torch.tensor(max_norm > 0), (max_norm / (norm + 1e-6)).clamp(max=1.0), torch.ones_like(norm))
    gc = g * coef[:, None]
    return w - lr * (gc + wd * w), gc


def test_fused_client_step_matches_clip_stats_sgd_cpu():
    w, g = _rows()
    hyper = arena_ops.make_hyper(3, "cpu", lr=0.1, max_norm=5.0, weight_decay=0.01)
    stats = torch.zeros(3, 4)
    w_ref, gc = _manual_step(w.clone(), g.clone(), 0.1, 5.0, 0.01)
    arena_ops.fused_client_step(w, g, hyper, stats, n_logical=4000)
    assert torch.allclose(w, w_ref, atol=1e-6)
    assert torch.all(g == 0)
    assert torch.allclose(stats[:, 0], gc.sum(1), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[:, 1], (gc * gc).sum(1), rtol=1e-5)
    assert torch.all(stats[:, 2] == 4000)
    mean, mag, var, norm = arena_ops.finalize_stats(stats)
    assert torch.allclose(norm, gc.norm(dim=1), rtol=1e-5)
    assert torch.allclose(var, torch.zeros(3), atol=1e-5)      # the reference's var = E[g^2] - mag^2 == 0


def test_fused_client_step_equals_torch_sgd_momentum():
    torch.manual_seed(0)
    lin = torch.nn.Linear(37, 11)
    ref = torch.nn.Linear(37, 11)
    ref.load_state_dict(lin.state_dict())
    opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=True)
    wa, ga = adopt_module(lin)
    mom = torch.zeros_like(wa.flat)
    first = torch.ones(1, dtype=torch.int32)
    hyper = arena_ops.make_hyper(1, "cpu", lr=0.05, max_norm=0.0, weight_decay=1e-3, momentum=0.9)
    stats = torch.zeros(1, 4)
    for step in range(4):
        x = torch.randn(5, 37)
        for m in (lin, ref):
            m.zero_grad() if m is ref else None
        ref(x).pow(2).sum().backward()
        lin(x).pow(2).sum().backward()
        opt.step(); opt.zero_grad()
        arena_ops.fused_client_step(wa.flat, ga.flat, hyper, stats, mom, n_logical=wa.layout.numel, nesterov=True,
                                    first_step=first)
        first.zero_()
        for p, q in zip(lin.parameters(), ref.parameters()):
            assert torch.allclose(p, q, atol=1e-5), step


def test_accumulate_pseudo_grad_and_single():
    wl, _ = _rows(S=4, P=1024, seed=1)
    wg = torch.randn(1024)
    acc = torch.ones(1024)
    wts = torch.tensor([2.0, 0.5, 3.0, 7.0])
    active = torch.tensor([1, 1, 0, 1], dtype=torch.int32)
    expect = 1 + sum(w * a * (wg - wl[i]) for i, (w, a) in enumerate(zip(wts.tolist(), active.tolist())))
    arena_ops.accumulate_pseudo_grad(acc, wg, wl, wts, active)
    assert torch.allclose(acc, expect, atol=1e-5)
    out, st = torch.empty(1024), torch.zeros(4)
    arena_ops.pseudo_grad(wg, wl[0], out, torch.tensor(3.0), st)
    assert torch.allclose(out, 3 * (wg - wl[0]))
    assert math.isclose(st[1].item(), ((wg - wl[0]) ** 2).sum().item(), rel_tol=1e-5)


@pytest.mark.parametrize("kind", ["sgd", "adam", "adamW", "adamax", "lamb", "LarsSGD"])
def test_server_update_reference_matches_optimizer_classes(kind):
    """The flat-arena server update (what the CUDA kernel implements) == the per-tensor optimizer classes."""
    from msrflute_b200.utils import make_optimizer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(20, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    twin = torch.nn.Sequential(torch.nn.Linear(20, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    twin.load_state_dict(model.state_dict())
    cfg = {"type": kind, "lr": 0.01}
    if kind in ("sgd", "LarsSGD"):
        cfg["momentum"] = 0.9
    if kind in ("adam", "adamW", "adamax", "lamb"):
        cfg["weight_decay"] = 0.01
    if kind == "adamW":
        cfg["eps"] = 1e-6
    opt = make_optimizer(cfg, twin)
    wa, ga = adopt_module(model)
    g0 = opt.param_groups[0]
    st = arena_ops.ServerOptState(kind, wa.flat.numel(), "cpu", lr=0.01, betas=g0.get("betas", (0.9, 0.999)),
                                  eps=g0.get("eps", 1e-8), weight_decay=g0.get("weight_decay", 0.0),
                                  momentum=g0.get("momentum", 0.0), correct_bias=g0.get("correct_bias", True))
    segs = wa.layout.segments()
    for it in range(3):
        grads = [torch.randn_like(p) for p in twin.parameters()]
        for p, g in zip(twin.parameters(), grads):
            p.grad = g.clone()
        opt.step()
        acc = torch.zeros_like(wa.flat)
        for v, g in zip(wa.layout.views(acc), grads):
            v.copy_(g * 4.0)                               # two "ranks" with weight sum 8
        arena_ops.server_update(wa.flat, [acc, acc.clone()], torch.tensor(8.0), st, segments=segs)
        for p, q in zip(model.parameters(), twin.parameters()):
            assert torch.allclose(p, q, atol=2e-5), (kind, it, (p - q).abs().max())


def test_arena_layout_alignment_and_views():
    lay = ArenaLayout([torch.Size([3, 5]), torch.Size([7]), torch.Size([])])
    assert all(o % 32 == 0 for o in lay.offsets) and lay.padded_numel % 32 == 0 and lay.numel == 23
    m = torch.nn.Linear(5, 3)
    w, g = adopt_module(m)
    assert module_arena(m) is not None
    m.weight.data.fill_(2.0)
    assert w.flat[:15].eq(2.0).all() and w.flat[15:32].eq(0).all()
    m(torch.randn(2, 5)).sum().backward()
    assert g.flat[:15].abs().sum() > 0


def test_compact_slot_plan_and_mapped_ops_match_full_arena():
    """Dead-tap elision: the compact slot layout + mapped broadcast/gather are equivalent to the full arenas."""
    import torch
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slot_resnet import SlotBatchedResNet
    from msrflute_b200.ops import arena_ops
    from msrflute_b200.ops.slot_ops import live_taps
    from msrflute_b200.parallel.arena import ArenaLayout

    assert live_taps(1, 1, 3, 3, 1, 1) == [(1, 1)]
    assert live_taps(2, 2, 3, 3, 2, 1) == [(1, 1), (1, 2), (2, 1), (2, 2)]
    assert len(live_taps(8, 8, 3, 3, 1, 1)) == 9
    torch.manual_seed(0)
    model = RESNET({"group_norm": 2, "num_classes": 10})
    lay = ArenaLayout.from_module(model)
    plan = SlotBatchedResNet.plan_compact(model, lay, torch.zeros(2, 3, 32, 32))
    assert plan is not None and plan["numel"] < 0.45 * lay.numel            # layer4's 3x3 filters keep 1 tap of 9
    assert any("layer4" in n for n in plan["compact"]) and not any("layer1" in n for n in plan["compact"])
    m = plan["index_map"]
    live = m[m >= 0].long()
    assert live.unique().numel() == live.numel() and int(live.max()) < lay.padded_numel
    S, P = 3, lay.padded_numel
    wg = torch.randn(P)
    W = torch.zeros(S, plan["numel"])
    arena_ops.scatter_in(W, wg, m)
    assert torch.equal(W[1][m >= 0], wg[live]) and float(W[:, m < 0].abs().sum()) == 0
    W += torch.randn_like(W) * (m >= 0)
    Wfull = wg.view(1, -1).repeat(S, 1)
    Wfull[:, live] = W[:, m >= 0]
    wts, act = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1, 0, 1], dtype=torch.int32)
    acc_c, acc_f = torch.zeros(P), torch.zeros(P)
    arena_ops.accumulate_pseudo_grad_mapped(acc_c, wg, W, wts, act, m)
    arena_ops.accumulate_pseudo_grad(acc_f, wg, Wfull, wts, act)
    assert torch.allclose(acc_c, acc_f, atol=1e-5)
