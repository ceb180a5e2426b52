"""CUDA kernels vs their PyTorch fp32 references (run on the B200 box: ``pytest -m gpu``)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ext():
    from msrflute_b200.ops import _ext
    return _ext.load(required=True)


def _cpu_ref(fn, *tensors, **kw):
    """Run an arena op on CPU copies (the PyTorch reference path) and return the copies."""
    cp = [t.detach().cpu().clone() if torch.is_tensor(t) else t for t in tensors]
    kw = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in kw.items()}
    fn(*cp, **kw)
    return cp


@pytest.mark.parametrize("S,P", [(1, 128), (3, 4096 + 32), (10, 1 << 20), (2, 11_700_000 // 32 * 32)])
@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_fused_client_step(S, P, momentum):
    from msrflute_b200.ops import arena_ops
    _ext()
    torch.manual_seed(0)
    w = torch.randn(S, P, device="cuda")
    g = torch.randn(S, P, device="cuda") * 2
    mom = torch.randn(S, P, device="cuda") if momentum else None
    first = torch.zeros(S, dtype=torch.int32, device="cuda") if momentum else None
    if first is not None:
        first[0] = 1
    hyper = arena_ops.make_hyper(S, "cuda", lr=0.1, max_norm=5.0, weight_decay=0.01, momentum=momentum)
    hyper[-1, 1] = 0.0                                            # last row: clipping disabled
    stats = torch.rand(S, 4, device="cuda")
    ref = _cpu_ref(arena_ops.fused_client_step, w, g, hyper, stats, mom, n_logical=P - 7, nesterov=bool(momentum),
                   first_step=first)
    arena_ops.fused_client_step(w, g, hyper, stats, mom, n_logical=P - 7, nesterov=bool(momentum), first_step=first)
    torch.cuda.synchronize()
    assert torch.allclose(w.cpu(), ref[0], atol=1e-5, rtol=1e-5)
    assert torch.all(g == 0)
    assert torch.allclose(stats.cpu(), ref[3], rtol=2e-4, atol=1e-2)
    if mom is not None:
        assert torch.allclose(mom.cpu(), ref[4], atol=1e-5, rtol=1e-5)


def test_clip_and_stats_and_pseudo_grad_and_accumulate():
    from msrflute_b200.ops import arena_ops
    _ext()
    torch.manual_seed(1)
    S, P = 5, 65536 + 64
    g = torch.randn(S, P, device="cuda") * 4
    hyper = arena_ops.make_hyper(S, "cuda", lr=0.0, max_norm=3.0)
    stats = torch.zeros(S, 4, device="cuda")
    ref = _cpu_ref(arena_ops.clip_and_stats, g, hyper, stats, n_logical=P)
    arena_ops.clip_and_stats(g, hyper, stats, n_logical=P)
    assert torch.allclose(g.cpu(), ref[0], atol=1e-6, rtol=1e-5)
    assert torch.allclose(stats.cpu(), ref[2], rtol=2e-4, atol=1e-2)

    wg, wl = torch.randn(P, device="cuda"), torch.randn(S, P, device="cuda")
    out, st = torch.empty(P, device="cuda"), torch.zeros(4, device="cuda")
    arena_ops.pseudo_grad(wg, wl[1], out, torch.tensor(2.5, device="cuda"), st)
    assert torch.allclose(out, 2.5 * (wg - wl[1]), atol=1e-6)
    assert math.isclose(st[1].item(), ((wg - wl[1]) ** 2).sum().item(), rel_tol=1e-4)

    acc = torch.randn(P, device="cuda")
    wts = torch.tensor([1.0, 100.0, 0.0, 3.5, 20.0], device="cuda")
    act = torch.tensor([1, 1, 1, 0, 1], dtype=torch.int32, device="cuda")
    expect = acc + ((wg[None] - wl) * (wts * act)[:, None]).sum(0)
    arena_ops.accumulate_pseudo_grad(acc, wg, wl, wts, act)
    assert torch.allclose(acc, expect, atol=1e-3, rtol=1e-5)


@pytest.mark.parametrize("kind", ["sgd", "adam", "adamW", "adamax", "lamb", "LarsSGD"])
@pytest.mark.parametrize("clip", [None, 0.5])
def test_server_update_kernel(kind, clip):
    from msrflute_b200.ops import arena_ops
    from msrflute_b200.parallel.arena import ArenaLayout
    _ext()
    torch.manual_seed(2)
    lay = ArenaLayout([torch.Size([64, 33]), torch.Size([64]), torch.Size([10, 64]), torch.Size([10])])
    P = lay.padded_numel
    mask = torch.zeros(P)
    for v in lay.views(mask):
        v.fill_(1.0)
    w0 = torch.randn(P) * mask
    kw = dict(lr=0.01, weight_decay=0.01 if kind != "sgd" else 0.0, momentum=0.9 if kind in ("sgd", "LarsSGD") else 0.0,
              eps=1e-6 if kind in ("adamW", "lamb") else 1e-8)
    st_gpu = arena_ops.ServerOptState(kind, P, "cuda", **kw)
    st_cpu = arena_ops.ServerOptState(kind, P, "cpu", **kw)
    w_gpu, w_cpu = w0.cuda(), w0.clone()
    peer = torch.zeros(P, device="cuda")
    for it in range(3):
        accs = [torch.randn(P) * mask for _ in range(3)]
        wsum = torch.tensor(7.0)
        stats = torch.zeros(2, device="cuda")
        arena_ops.server_update(w_gpu, [a.cuda() for a in accs], wsum.cuda(), st_gpu, max_grad_norm=clip,
                                segments=lay.segments("cuda"), bcast=[w_gpu, peer], stats_out=stats if clip else None)
        arena_ops.server_update(w_cpu, [a.clone() for a in accs], wsum, st_cpu, max_grad_norm=clip,
                                segments=lay.segments())
        assert torch.allclose(w_gpu.cpu(), w_cpu, atol=3e-5, rtol=1e-4), (kind, it, (w_gpu.cpu() - w_cpu).abs().max())
        assert torch.equal(peer, w_gpu)                            # fused broadcast into the peer buffer
        if clip:
            g = sum(accs) / 7.0
            assert math.isclose(stats[0].item(), g.norm().item(), rel_tol=1e-4)


def test_server_update_dp_noise_statistics_and_invariance():
    from msrflute_b200.ops import arena_ops
    _ext()
    P = 1 << 20
    st = arena_ops.ServerOptState("sgd", P, "cuda", lr=1.0)
    outs = []
    for split in (1, 3):                                         # same total, different number of "ranks"
        w = torch.zeros(P, device="cuda")
        accs = [torch.zeros(P, device="cuda") for _ in range(split)]
        arena_ops.server_update(w, accs, torch.tensor(1.0, device="cuda"), st, noise_scale=0.25, seed=1234)
        outs.append(w.clone())
    assert torch.equal(outs[0], outs[1])                         # Philox keyed by element index only
    z = -outs[0] / 0.25
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1.0) < 5e-3
    w2 = torch.zeros(P, device="cuda")
    arena_ops.server_update(w2, [torch.zeros(P, device="cuda")], torch.tensor(1.0, device="cuda"), st, noise_scale=0.25,
                            seed=99)
    assert not torch.equal(w2, outs[0])


@pytest.mark.parametrize("shape,cpg", [((20, 64, 16, 16), 2), ((20, 128, 4, 4), 2), ((20, 512, 1, 1), 2), ((4, 32, 7, 5), 8)])
@pytest.mark.parametrize("per_group", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fuse", [(False, False), (True, True)])
def test_group_norm_fwd_bwd(shape, cpg, per_group, dtype, fuse):
    from msrflute_b200.ops import norm_ops
    _ext()
    use_res, relu = fuse
    torch.manual_seed(3)
    N, C = shape[:2]
    G = C // cpg
    x = (torch.randn(shape, device="cuda") * 2 + 0.5).to(dtype).requires_grad_(True)
    res = torch.randn(shape, device="cuda").to(dtype).requires_grad_(True) if use_res else None
    A = G if per_group else C
    wt = (torch.rand(A, device="cuda") + 0.5).requires_grad_(True)
    b = torch.randn(A, device="cuda").requires_grad_(True)
    y = norm_ops.group_norm(x, G, wt, b, 1e-5, residual=res, relu=relu, per_group_affine=per_group)
    dy = torch.randn_like(y)
    y.backward(dy)
    got = [y.detach().float(), x.grad.float(), wt.grad.clone(), b.grad.clone()] + ([res.grad.float()] if use_res else [])
    x2 = x.detach().float().requires_grad_(True)
    r2 = res.detach().float().requires_grad_(True) if use_res else None
    w2, b2 = wt.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y2 = norm_ops._group_norm_ref(x2, G, w2, b2, 1e-5, r2, relu, per_group)
    y2.backward(dy.float())
    want = [y2.detach(), x2.grad, w2.grad, b2.grad] + ([r2.grad] if use_res else [])
    tol = dict(atol=2e-4, rtol=2e-4) if dtype == torch.float32 else dict(atol=6e-2, rtol=6e-2)
    for a, e in zip(got, want):
        assert torch.allclose(a, e, **tol), (a - e).abs().max()


@pytest.mark.parametrize("path", ["fma", "tcgen05-r1", "slotnet"])
def test_engine_round_matches_generic_path(path, monkeypatch):
    """One FedAvg round through the device engine (graphs, slots, fused accumulate) == the generic per-client path.
    ``fma``: the round-1 slot executor with exact-fp32 convolutions, tight tolerance.  ``tcgen05-r1`` / ``slotnet``: tf32
    tensor-core convolutions (cuDNN's default precision too); GroupNorm over 2-channel groups on 1x1 feature maps is a
    sign function of a difference, so 1e-3 operand rounding is amplified ~1e4x in this configuration — only a sanity
    bound is asserted here; the tf32 kernels are checked against fp32 with tight bounds in
    ``test_slotnet_gemm_gpu.py`` / ``test_slot_conv_matches_conv2d`` and the training trajectory in
    ``test_flagship_tf32_trajectory_tracks_fp32``."""
    import bench
    from msrflute_b200.ops import slot_ops
    monkeypatch.setenv("FLUTE_SLOTNET", "1" if path == "slotnet" else "0")
    slot_ops.set_conv_impl("fma" if path == "fma" else "auto")
    try:
        _engine_round_check(bench, 0.02 if path == "fma" else 0.5)
    finally:
        slot_ops.set_conv_impl("auto")


def _loss_trajectory(monkeypatch, slotnet, impl, rounds):
    import bench
    from msrflute_b200.ops import slot_ops
    monkeypatch.setenv("FLUTE_SLOTNET", "1" if slotnet else "0")
    slot_ops.set_conv_impl(impl)
    torch.manual_seed(1234)
    job = bench.build_flagship(n_clients_per_round=4, users=8, norm="gn")
    import random
    random.seed(99)
    torch.manual_seed(99)
    try:
        return [job.run_round() for _ in range(rounds)]
    finally:
        job.close()
        slot_ops.set_conv_impl("auto")


def test_flagship_tf32_trajectory_tracks_fp32(monkeypatch):
    """The benchmarked numerics: 10 rounds of the flagship config (GroupNorm 2 ch/group down to 1x1 maps) with the tf32
    SlotNet executor track the exact-fp32 executor's training-loss trajectory.  Same seeds => same clients and the
    same shuffles; individual gradients differ (see above) but the optimisation must not."""
    ref = _loss_trajectory(monkeypatch, slotnet=False, impl="fma", rounds=10)
    got = _loss_trajectory(monkeypatch, slotnet=True, impl="auto", rounds=10)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "trajectory_tf32_vs_fp32.txt"), "w") as f:
        f.write("round fp32_exact tf32_slotnet rel_diff\n")
        for i, (a, b) in enumerate(zip(ref, got)):
            f.write("{} {:.5f} {:.5f} {:.4f}\n".format(i, a, b, abs(a - b) / abs(a)))
    assert all(math.isfinite(v) for v in got)
    assert abs(got[0] - ref[0]) / ref[0] < 0.01, (ref[0], got[0])          # same weights, first round: ~tf32 noise
    for a, b in zip(ref, got):
        assert abs(a - b) / a < 0.08, (ref, got)
    assert min(got[5:]) < got[0] and min(ref[5:]) < ref[0], (ref, got)


def _engine_round_check(bench, tol):
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False          # the generic path uses cuDNN: compare fp32 with fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    job = bench.build_flagship(n_clients_per_round=3, users=12, norm="gn")
    srv, worker = job.server, job.worker
    assert worker.engine is not None
    # one full-batch step per client => the result does not depend on the (different) shuffles of the two paths
    job.config["client_config"]["data_config"]["train"]["batch_size"] = 100
    srv.begin_training()
    from msrflute_b200.parallel.arena import module_arena
    w0 = module_arena(srv.worker_trainer.model)[0].flat.clone()
    ids = [0, 5, 7]
    worker.set_weights(w0)
    worker.accumulator().zero_()
    outs_e = worker.engine.train_clients(ids, 0.1, 0, worker.weight_buffer(), worker.accumulator())
    acc_e = worker.accumulator().clone()
    worker.accumulator().zero_()
    eng, worker.engine = worker.engine, None
    outs_g = worker.train_clients(ids, (0.1, None, 0), fused=True)
    acc_g = worker.accumulator().clone()
    worker.engine = eng
    assert [o["ns"] for o in outs_e] == [o["ns"] for o in outs_g] == [100, 100, 100]
    assert [o["pl"]["weight"] for o in outs_e] == [100.0] * 3
    ne, ng = acc_e.norm().item(), acc_g.norm().item()
    assert abs(ne / ng - 1.0) < tol, (ne, ng)
    cos = torch.dot(acc_e, acc_g) / (acc_e.norm() * acc_g.norm())
    assert cos > 1.0 - tol, cos
    for a, b in zip(outs_e, outs_g):
        assert abs(a["tl"] / b["tl"] - 1.0) < tol / 2 and abs(float(a["rg"]) / float(b["rg"]) - 1.0) < tol / 2
    for o in outs_e:
        assert o["tl"] > 0 and math.isfinite(o["tl"]) and o["rg"] > 0
    srv.end_training()


def test_flagship_rounds_reduce_loss():
    import bench
    job = bench.build_flagship(n_clients_per_round=4, users=8, norm="gn")
    losses = [job.run_round() for _ in range(6)]
    job.close()
    assert all(math.isfinite(l) for l in losses)
    assert min(losses[3:]) < losses[0], losses


@pytest.mark.parametrize("G,M,N,K", [(1, 128, 64, 64), (1, 20, 1000, 512), (1, 1280, 64, 576), (1, 320, 90, 256),
                                     (3, 200, 130, 200), (1, 4096, 4096, 1024), (10, 80, 256, 2304)])
@pytest.mark.parametrize("epi", [(False, False, True), (True, True, False)])
@pytest.mark.parametrize("impl", [0, 1, 2, 3])      # auto | one tile per CTA | persistent BLOCK_N 128 | persistent 256
def test_gemm_tcgen05_matches_fp32_reference(G, M, N, K, epi, impl):
    ext = _ext()
    ext.gemm_set_impl(impl)
    try:
        _gemm_check(G, M, N, K, epi)
    finally:
        ext.gemm_set_impl(0)


def _gemm_check(G, M, N, K, epi):
    ext = _ext()
    if not hasattr(ext, "gemm_bf16_tn"):
        pytest.skip("GEMM not built")
    use_bias, relu, out_fp32 = epi
    torch.manual_seed(4)
    a = torch.randn(G, M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(G, N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") if use_bias else None
    if G == 1:
        c = ext.gemm_bf16_tn(a[0].contiguous(), b[0].contiguous(), bias, relu, out_fp32)[None]
    else:
        c = ext.gemm_bf16_tn(a, b, bias, relu, out_fp32)
    ref = torch.einsum("gmk,gnk->gmn", a.float(), b.float())
    if use_bias:
        ref = ref + bias
    if relu:
        ref = ref.relu()
    err = (c.float() - ref).abs().max().item()
    # fp32 out: only bf16-input products are exact, accumulation order differs; bf16 out: + 2^-8 relative rounding
    tol = 2e-3 * math.sqrt(K) if out_fp32 else 2e-3 * math.sqrt(K) + ref.abs().max().item() * 2 ** -8
    assert err < tol, (err, tol)
    assert c.dtype == (torch.float32 if out_fp32 else torch.bfloat16)


@pytest.mark.parametrize("cfg", [  # (S, B, Cin, H, Cout, k, stride, pad)
    (3, 20, 3, 32, 64, 7, 2, 3), (2, 20, 64, 8, 64, 3, 1, 1), (4, 20, 64, 8, 128, 3, 2, 1), (2, 20, 64, 8, 128, 1, 2, 0),
    (3, 20, 256, 2, 512, 3, 2, 1), (2, 20, 512, 1, 512, 3, 1, 1), (1, 5, 7, 9, 11, 3, 1, 1),
    (2, 20, 128, 4, 128, 3, 1, 1), (2, 20, 256, 2, 256, 3, 1, 1), (2, 7, 5, 13, 70, 5, 2, 2)])
@pytest.mark.parametrize("impl", ["fma", "tcgen05"])
def test_slot_conv_matches_conv2d(cfg, impl):
    """Both implementations of the slot-batched convolution (fp32 FMA tiles; tcgen05 kind::tf32 implicit GEMM with
    producer-warp im2col) against per-slot ``F.conv2d`` in fp32."""
    ext = _ext()
    if not hasattr(ext, "slot_conv_fprop"):
        pytest.skip("slot conv not built")
    ext.slot_conv_set_impl(1 if impl == "fma" else 2)
    try:
        _check_slot_conv(cfg, tf32=impl == "tcgen05")
    finally:
        ext.slot_conv_set_impl(0)


def _check_slot_conv(cfg, tf32):
    from msrflute_b200.ops.slot_ops import SlotConv2d
    S, B, Cin, H, Cout, k, stride, pad = cfg
    torch.manual_seed(5)
    n = Cout * Cin * k * k
    P = ((n + 64 + 31) // 32) * 32
    off = 32
    W = torch.zeros(S, P, device="cuda")
    W[:, off:off + n] = torch.randn(S, n, device="cuda") * 0.1
    G = torch.zeros(S, P, device="cuda")
    x = torch.randn(S, B, Cin, H, H, device="cuda", requires_grad=True)
    dummy = torch.zeros(1, device="cuda", requires_grad=True)
    y = SlotConv2d.apply(x, dummy, W, G, off, Cout, k, k, stride, pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.backends.cudnn.allow_tf32 = False
    for s in range(S):
        # reference in float64: cuDNN's fp32 algorithm choice (FFT / Winograd variants) is itself only ~1e-3 accurate
        # for some shapes, which made the exact-fp32 comparison depend on its autotuner
        w64 = W[s, off:off + n].view(Cout, Cin, k, k).double().requires_grad_(True)
        xs64 = x[s].detach().double().requires_grad_(True)
        ys64 = torch.nn.functional.conv2d(xs64, w64, None, stride, pad)
        ys64.backward(dy[s].double())
        ys, w, xs = ys64.float(), w64, xs64

        wg, xg = w64.grad.float(), xs64.grad.float()
        gw = G[s, off:off + n].view(Cout, Cin, k, k)
        if tf32:        # 10-bit mantissa operands, fp32 accumulation: bound the error by the size of the result
            for got, ref in ((y[s], ys), (x.grad[s], xg), (gw, wg)):
                err = (got - ref).abs().max().item()
                assert err <= 4e-3 * ref.abs().max().item() + 1e-5, (err, ref.abs().max().item())
            continue
        assert torch.allclose(y[s], ys, atol=2e-3, rtol=1e-3), (y[s] - ys).abs().max()
        assert torch.allclose(x.grad[s], xg, atol=2e-3, rtol=1e-3), (x.grad[s] - xg).abs().max()
        assert torch.allclose(gw, wg, atol=5e-3, rtol=2e-3), (gw - wg).abs().max()
    assert G[:, :off].abs().sum() == 0 and G[:, off + n:].abs().sum() == 0         # nothing written outside the tensor


@pytest.mark.parametrize("conv_impl", ["fma", "auto"])
def test_slot_batched_resnet_matches_per_client_models(conv_impl):
    ext = _ext()
    if not hasattr(ext, "slot_conv_fprop"):
        pytest.skip("slot conv not built")
    from msrflute_b200.ops import slot_ops
    slot_ops.set_conv_impl(conv_impl)
    try:
        if conv_impl == "fma":      # the flagship configuration (2 channels per group, 32x32 inputs), exact fp32 convs
            _slot_resnet_check(loss_tol=2e-3, grad_tol=2.5e-2, cpg=2, hw=32)
        else:                       # tf32 convs: a configuration where GroupNorm does not amplify rounding chaotically
            _slot_resnet_check(loss_tol=1e-2, grad_tol=5e-2, cpg=32, hw=32)
    finally:
        slot_ops.set_conv_impl("auto")


def _slot_resnet_check(loss_tol, grad_tol, cpg, hw):
    import copy
    from msrflute_b200.models.resnet_gn import RESNET
    from msrflute_b200.models.slot_resnet import SlotBatchedResNet
    from msrflute_b200.parallel.arena import ArenaLayout, adopt_module
    torch.manual_seed(6)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    S, B = 3, 20
    base = RESNET({"group_norm": cpg, "num_classes": 100}).cuda()
    assert SlotBatchedResNet.supports(base)
    lay = ArenaLayout.from_module(base)
    W = torch.zeros(S, lay.padded_numel, device="cuda")
    G = torch.zeros_like(W)
    models = []
    for s in range(S):
        m = copy.deepcopy(base)
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        adopt_module(m, with_grad=True, param_buffer=W[s], grad_buffer=torch.zeros(lay.padded_numel, device="cuda"))
        models.append(m)
    slot = SlotBatchedResNet(models[0], lay, W, G)
    x = torch.randn(S, B, 3, hw, hw, device="cuda") * 50 + 100
    y = torch.randint(0, 100, (S, B), device="cuda")
    losses = slot.losses(x, y)
    losses.sum().backward()
    for s in range(S):
        m = models[s]
        m.zero_grad()
        l = m.loss({"x": x[s], "y": y[s]})
        l.backward()
        assert abs(l.item() - losses[s].item()) < loss_tol * max(1.0, abs(l.item())), (l.item(), losses[s].item())
        ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        got = torch.cat([v.reshape(-1) for v in lay.views(G[s])])
        rel = (ref - got).norm() / ref.norm()
        assert rel < grad_tol, rel     # fp32 atomics + GroupNorm over 2-element groups amplify rounding


def test_tc_linear_forward_backward_matches_fp32():
    ext = _ext()
    from msrflute_b200.ops.linear_ops import TCLinear, swap_linear_modules
    torch.manual_seed(7)
    lin = TCLinear(768, 3072).cuda()
    ref = torch.nn.Linear(768, 3072).cuda()
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(4, 64, 768, device="cuda", requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y = lin(x)
    yr = ref(xr)
    assert y.dtype == torch.bfloat16 and (y.float() - yr).abs().max() < 0.06
    dy = torch.randn_like(yr)
    y.backward(dy.to(y.dtype))
    yr.backward(dy)
    assert (x.grad - xr.grad).abs().max() / xr.grad.abs().max() < 0.03
    assert (lin.weight.grad - ref.weight.grad).abs().max() / ref.weight.grad.abs().max() < 0.03
    assert (lin.bias.grad - ref.bias.grad).abs().max() / ref.bias.grad.abs().max() < 0.03
    mlp = torch.nn.Sequential(torch.nn.Linear(128, 256), torch.nn.ReLU(), torch.nn.Linear(256, 16)).cuda()
    assert swap_linear_modules(mlp, min_features=64) == 1          # the 256→16 head is below the size threshold


def test_rnn_cells_match_reference():
    ext = _ext()
    if not hasattr(ext, "gru_cell_fwd"):
        pytest.skip("rnn kernels not built")
    from msrflute_b200.ops import rnn_ops
    torch.manual_seed(8)
    B, H = 7, 96
    gi, gh, h = (torch.randn(B, 3 * H, device="cuda", requires_grad=True), torch.randn(B, 3 * H, device="cuda", requires_grad=True),
                 torch.randn(B, H, device="cuda", requires_grad=True))
    out = rnn_ops.gru_cell(gi, gh, h)
    ref = rnn_ops._gru_cell_ref(gi, gh, h)
    assert torch.allclose(out, ref, atol=1e-5)
    g = torch.randn_like(out)
    got = torch.autograd.grad(out, (gi, gh, h), g)
    want = torch.autograd.grad(ref, (gi, gh, h), g)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-4), (a - b).abs().max()
    gates, c = torch.randn(B, 4 * H, device="cuda", requires_grad=True), torch.randn(B, H, device="cuda", requires_grad=True)
    h2, c2 = rnn_ops.lstm_cell(gates, c)
    hr, cr = rnn_ops._lstm_cell_ref(gates, c)
    assert torch.allclose(h2, hr, atol=1e-5) and torch.allclose(c2, cr, atol=1e-5)
    gh_, gc_ = torch.randn_like(h2), torch.randn_like(c2)
    got = torch.autograd.grad((h2, c2), (gates, c), (gh_, gc_))
    want = torch.autograd.grad((hr, cr), (gates, c), (gh_, gc_))
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-4), (a - b).abs().max()


# ------------------------------------------------------------------------------------------------ misc kernels (K7/K14/K15/K19)
def test_quantize_segments_kernel_matches_reference_transform():
    _ext()
    from msrflute_b200.ops import quant_ops
    from msrflute_b200.extensions.quantization.quant import quantize_tensor_
    torch.manual_seed(11)
    segs = [(0, 1000), (1024, 37), (1088, 70000), (71104, 1)]          # gaps = alignment padding, must stay untouched
    flat = torch.randn(71200, device="cuda")
    flat[1000:1024] = 7.0
    ref = flat.clone()
    for o, n in segs:
        quantize_tensor_(ref[o:o + n], 6, 0.7)
    got = quant_ops.quantize_segments_(flat.clone(), segs, 6, 0.7)
    assert torch.equal(got[1000:1024], ref[1000:1024])                  # padding untouched
    mism = (got - ref).abs() > 1e-5
    assert mism.float().mean().item() < 1e-4, mism.float().mean()       # (ties at a level boundary may round differently)
    for o, n in segs[:3]:
        assert abs((got[o:o + n] == 0).float().mean().item() - (ref[o:o + n] == 0).float().mean().item()) < 2e-3


def test_local_dp_kernel_clip_and_noise():
    _ext()
    from msrflute_b200.ops import misc_ops
    torch.manual_seed(12)
    g = torch.randn(1_000_003, device="cuda") * 3
    n0 = g.norm().item()
    h = g.clone()
    norm = misc_ops.local_dp_(h, 2.0, 0.0, True)
    assert abs(norm.item() - n0) < 1e-2 * n0 and abs(h.norm().item() - 2.0) < 1e-3
    small = torch.randn(1000, device="cuda") * 1e-3
    keep = small.clone()
    misc_ops.local_dp_(small, 2.0, 0.0, True)
    assert torch.allclose(small, keep)                                   # below the bound: untouched
    h = g.clone()
    misc_ops.local_dp_(h, 2.0, 0.5, False, seed=1234)
    noise = h - g * (2.0 / n0)
    assert abs(noise.mean().item()) < 5e-3 and abs(noise.std().item() - 0.5) < 5e-3
    h2 = g.clone()
    misc_ops.local_dp_(h2, 2.0, 0.5, False, seed=1234)
    assert (h - h2).abs().max().item() < 1e-5                            # counter-based noise: reproducible (the norm's
                                                                         # atomic reduction order may change the last bit)
    h3 = g.clone()
    misc_ops.local_dp_(h3, 2.0, 0.5, False, seed=1235)
    assert (h - h3).abs().max().item() > 0.1


@pytest.mark.parametrize("shape", [(200, 1000), (7, 10), (33, 62), (64, 4097)])
def test_softmax_ce_kernel_matches_torch(shape):
    _ext()
    from msrflute_b200.ops import misc_ops
    torch.manual_seed(13)
    rows, C = shape
    x = (torch.randn(rows, C, device="cuda") * 4).requires_grad_(True)
    t = torch.randint(0, C, (rows,), device="cuda")
    t[0] = -100
    wts = torch.rand(rows, device="cuda")
    loss = misc_ops.softmax_cross_entropy(x, t)
    (loss * wts).sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(x2, t, reduction="none", ignore_index=-100)
    (ref * wts).sum().backward()
    assert torch.allclose(loss, ref, atol=2e-5, rtol=1e-5), (loss - ref).abs().max()
    assert torch.allclose(x.grad, x2.grad, atol=2e-6, rtol=1e-4), (x.grad - x2.grad).abs().max()


def test_cosine_stats_kernel():
    _ext()
    from msrflute_b200.ops import misc_ops
    torch.manual_seed(14)
    a, b = torch.randn(3_000_001, device="cuda"), torch.randn(3_000_001, device="cuda")
    s = misc_ops.cosine_stats(a, b)
    ref = torch.stack([torch.dot(a.double(), b.double()), torch.dot(a.double(), a.double()), torch.dot(b.double(), b.double())]).float()
    assert torch.allclose(s, ref, rtol=1e-4, atol=1.0), (s, ref)
    assert abs(misc_ops.cosine(a, 2 * a).item() - 1.0) < 1e-5
    from msrflute_b200.utils import compute_grad_cosines
    cs = compute_grad_cosines([[a[:100], a[100:]], [b]], [a])
    assert abs(cs[0] - 1.0) < 1e-5 and abs(cs[1] - (ref[0] / (ref[1] * ref[2]).sqrt()).item()) < 1e-4


@pytest.mark.parametrize("cfg", [(2, 20, 512, 1, 512, 3, 1, 1), (3, 20, 256, 2, 512, 3, 2, 1), (2, 20, 64, 1, 96, 5, 1, 2)])
@pytest.mark.parametrize("impl", ["fma", "tcgen05"])
def test_slot_conv_compact_weights_match_full_layout(cfg, impl):
    """Compact slot arenas store only the live taps of a filter ([Cout, Cin, ntaps]); outputs and input gradients must
    equal the full-layout run and the weight gradient must equal the live taps of the full-layout gradient."""
    ext = _ext()
    from msrflute_b200.ops.slot_ops import SlotConv2d, live_taps
    S, B, Cin, H, Cout, k, stride, pad = cfg
    taps = live_taps(H, H, k, k, stride, pad)
    assert len(taps) < k * k
    torch.manual_seed(21)
    ext.slot_conv_set_impl(1 if impl == "fma" else 2)
    try:
        n_full, n_c = Cout * Cin * k * k, Cout * Cin * len(taps)
        off = 64
        Wf = torch.zeros(S, off + n_full + 32, device="cuda")
        Wf[:, off:off + n_full] = torch.randn(S, n_full, device="cuda") * 0.1
        tap_idx = torch.tensor([kh * k + kw for kh, kw in taps], device="cuda")
        Wc = torch.zeros(S, off + n_c + 32, device="cuda")
        Wc[:, off:off + n_c] = Wf[:, off:off + n_full].view(S, Cout * Cin, k * k)[:, :, tap_idx].reshape(S, -1)
        Gf, Gc = torch.zeros_like(Wf), torch.zeros_like(Wc)
        x = torch.randn(S, B, Cin, H, H, device="cuda")
        dummy = torch.zeros(1, device="cuda", requires_grad=True)
        outs = []
        for W, G, compact in ((Wf, Gf, False), (Wc, Gc, True)):
            xi = x.clone().requires_grad_(True)
            y = SlotConv2d.apply(xi, dummy, W, G, off, Cout, k, k, stride, pad, compact)
            torch.manual_seed(22)
            y.backward(torch.randn_like(y))
            outs.append((y.detach(), xi.grad))
        tol = dict(atol=1e-4, rtol=1e-4) if impl == "fma" else dict(atol=2e-2, rtol=2e-2)
        assert torch.allclose(outs[0][0], outs[1][0], **tol) and torch.allclose(outs[0][1], outs[1][1], **tol)
        gf = Gf[:, off:off + n_full].view(S, Cout * Cin, k * k)
        assert torch.allclose(gf[:, :, tap_idx].reshape(S, -1), Gc[:, off:off + n_c], **tol)
        dead = torch.ones(k * k, dtype=torch.bool, device="cuda")
        dead[tap_idx] = False
        assert float(gf[:, :, dead].abs().sum()) == 0                  # dead taps really have zero gradient
        assert float(Gc[:, :off].abs().sum()) == 0 and float(Gc[:, off + n_c:].abs().sum()) == 0
    finally:
        ext.slot_conv_set_impl(0)


@pytest.mark.parametrize("shape,dtype,affine", [((64, 128, 768), torch.float32, True), ((5, 7, 100), torch.float32, True),
                                                ((1000, 256), torch.bfloat16, True), ((33, 48), torch.float32, False)])
def test_layer_norm_kernel_matches_torch(shape, dtype, affine):
    _ext()
    from msrflute_b200.ops import norm_ops
    torch.manual_seed(31)
    H = shape[-1]
    x = (torch.randn(shape, device="cuda") * 2 + 0.3).to(dtype).requires_grad_(True)
    w = (torch.rand(H, device="cuda") + 0.5).requires_grad_(True) if affine else None
    b = torch.randn(H, device="cuda").requires_grad_(True) if affine else None
    y = norm_ops.layer_norm(x, w, b, 1e-5)
    dy = torch.randn_like(y)
    y.backward(dy)
    x2 = x.detach().float().requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True) if affine else None
    b2 = b.detach().clone().requires_grad_(True) if affine else None
    y2 = torch.nn.functional.layer_norm(x2, (H,), w2, b2, 1e-5)
    y2.backward(dy.float())
    tol = dict(atol=2e-4, rtol=2e-4) if dtype == torch.float32 else dict(atol=6e-2, rtol=6e-2)
    assert torch.allclose(y.float(), y2, **tol), (y.float() - y2).abs().max()
    assert torch.allclose(x.grad.float(), x2.grad, **tol), (x.grad.float() - x2.grad).abs().max()
    if affine:
        rows = x.numel() // H
        assert torch.allclose(w.grad, w2.grad, atol=tol["atol"] * rows ** 0.5, rtol=tol["rtol"] * 5)
        assert torch.allclose(b.grad, b2.grad, atol=tol["atol"] * rows ** 0.5, rtol=tol["rtol"] * 5)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32)).cuda()
    assert norm_ops.swap_layer_norm_modules(m) == 1 and isinstance(m[1], norm_ops.FusedLayerNorm)
    assert torch.allclose(m(torch.ones(2, 16, device="cuda")), torch.nn.functional.layer_norm(
        m[0](torch.ones(2, 16, device="cuda")), (32,), m[1].weight, m[1].bias), atol=1e-5)


@pytest.mark.parametrize("shape,k,s,p", [((200, 64, 16, 16), 3, 2, 1), ((3, 5, 9, 13), 3, 2, 1), ((2, 4, 8, 8), 2, 2, 0),
                                         ((2, 3, 7, 7), 3, 1, 1)])
def test_max_pool2d_kernel_matches_torch(shape, k, s, p):
    _ext()
    from msrflute_b200.ops import misc_ops
    torch.manual_seed(41)
    x = torch.randn(shape, device="cuda").requires_grad_(True)
    y = misc_ops.max_pool2d(x, k, s, p)
    dy = torch.randn_like(y)
    y.backward(dy)
    x2 = x.detach().clone().requires_grad_(True)
    y2 = torch.nn.functional.max_pool2d(x2, k, s, p)
    y2.backward(dy)
    assert torch.equal(y, y2)
    assert torch.allclose(x.grad, x2.grad, atol=1e-6), (x.grad - x2.grad).abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(4096, 768, 3072), (512, 3072, 768), (200, 136, 264), (4096, 768, 768)])
def test_gemm_mn_major_operands_match_matmul(M, N, K):
    """C = A·Bᵀ with A and/or B handed over in their other storage order (what linear backward passes have)."""
    from msrflute_b200.ops import _ext
    C = _ext.load(required=True)
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)         # logical A [M, K]
    b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)         # logical B [N, K]
    want = a.float() @ b.float().t()
    at, bt = a.t().contiguous(), b.t().contiguous()                          # stored [K, M] / [K, N]
    for a_mn, b_mn in ((False, True), (True, False), (True, True)):
        got = C.gemm_bf16_mn(at if a_mn else a, bt if b_mn else b, a_mn, b_mn, True)
        err = float((got - want).norm() / want.norm())
        assert err < 2e-3, (a_mn, b_mn, err)


@pytest.mark.gpu
def test_tc_linear_backward_without_transposes_matches_fp32():
    from msrflute_b200.ops import linear_ops
    torch.manual_seed(1)
    x = (torch.randn(6, 100, 768, device="cuda") * 0.5).requires_grad_(True)
    w = (torch.randn(3072, 768, device="cuda") * 0.03).requires_grad_(True)
    b = torch.zeros(3072, device="cuda", requires_grad=True)
    y = linear_ops.linear(x, w, b)
    go = torch.randn_like(y.float()) * 0.1
    y.float().backward(go)
    g = [t.grad.clone() for t in (x, w, b)]
    for t in (x, w, b):
        t.grad = None
    (torch.nn.functional.linear(x, w, b)).backward(go)
    rel = lambda a_, b_: float((a_.float() - b_.float()).norm() / b_.float().norm())
    assert rel(g[0], x.grad) < 2e-2 and rel(g[1], w.grad) < 2e-2 and rel(g[2], b.grad) < 2e-2
