"""Device-engine coverage of FedProx and local differential privacy: the closed-form proximal
gradient in the fused client step and clip / normalise / Philox noise fused into the gather (K15 -> K21),
checked against the plain PyTorch formulation and against the generic per-client path."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_client_step_fedprox_matches_reference():
    from msrflute_b200.ops import arena_ops
    torch.manual_seed(0)
    S, P = 3, 4096
    w = torch.randn(S, P, device="cuda")
    g = torch.randn(S, P, device="cuda")
    ref = torch.randn(P, device="cuda")
    mult = torch.rand(P, device="cuda") * 0.3
    mult[-32:] = 0
    hyper = torch.tensor([[0.1, 5.0, 0.0, 0.0]] * S, device="cuda")
    stats, prox = torch.zeros(S, 4, device="cuda"), torch.zeros(S, device="cuda")
    w0, g0 = w.clone(), g.clone()
    arena_ops.fused_client_step(w, g, hyper, stats, None, n_logical=P, zero_grad=True, prox_ref=ref, prox_mult=mult,
                                prox_loss=prox)
    geff = g0 + mult * (w0 - ref)
    norm = geff.norm(dim=1, keepdim=True)
    coef = (5.0 / (norm + 1e-6)).clamp(max=1.0)
    want = w0 - 0.1 * coef * geff
    assert torch.allclose(w, want, atol=1e-5, rtol=1e-5)
    assert torch.allclose(prox, 0.5 * (mult * (w0 - ref) ** 2).sum(dim=1), rtol=1e-4)
    assert torch.allclose(stats[:, 3], norm.view(-1), rtol=1e-4)
    assert float(g.abs().max()) == 0.0


def test_slot_gather_kernels_match_reference():
    from msrflute_b200.ops import arena_ops
    torch.manual_seed(1)
    S, Pc, Pg = 4, 2048, 5000
    perm = torch.randperm(Pg, device="cuda")[:Pc].to(torch.int32)
    perm[::37] = -1
    wg = torch.randn(Pg, device="cuda")
    W = torch.zeros(S, Pc, device="cuda")
    row = torch.zeros(Pc, device="cuda")
    arena_ops.slot_gather_bcast(W, row, wg, perm)
    m = perm.long()
    want = torch.where(m >= 0, wg[m.clamp(min=0)], torch.zeros((), device="cuda"))
    assert torch.equal(row, want) and torch.equal(W, want.view(1, -1).expand(S, -1))
    W += torch.randn_like(W) * 0.1
    n2 = arena_ops.slot_pg_sqnorm(W, row, torch.zeros(S, device="cuda"))
    assert torch.allclose(n2, ((row - W) ** 2).sum(1), rtol=1e-4)
    coef = torch.tensor([1.0, 0.0, 2.5, 0.5], device="cuda")
    acc_slot = torch.zeros(Pc, device="cuda")
    arena_ops.slot_gather_fused(acc_slot, W, row, coef)
    assert torch.allclose(acc_slot, (coef.view(-1, 1) * (row - W)).sum(0), atol=1e-5, rtol=1e-5)
    # noise: N(0, sum sig^2) per coordinate, deterministic in the seeds
    sig = torch.tensor([0.5, 0.0, 1.0, 0.0], device="cuda")
    seeds = torch.tensor([11, 12, 13, 14], device="cuda")
    a1, a2 = torch.zeros(Pc, device="cuda"), torch.zeros(Pc, device="cuda")
    z = torch.zeros(S, device="cuda")
    arena_ops.slot_gather_fused(a1, W, row, z, sig, seeds)
    arena_ops.slot_gather_fused(a2, W, row, z, sig, seeds)
    assert torch.equal(a1, a2)
    assert abs(float(a1.std()) - math.sqrt(0.25 + 1.0)) < 0.08 and abs(float(a1.mean())) < 0.1
    acc = torch.zeros(Pg, device="cuda")
    acc_copy = acc_slot.clone()
    arena_ops.slot_scatter_acc(acc, acc_slot, perm)
    want_acc = torch.zeros(Pg, device="cuda")
    want_acc[m[m >= 0]] = acc_copy[m >= 0]
    assert torch.allclose(acc, want_acc) and float(acc_slot.abs().max()) == 0.0
    dead = torch.arange(100, 400, device="cuda", dtype=torch.int32)
    acc.zero_()
    arena_ops.dead_coord_noise(acc, dead, torch.tensor(4.0, device="cuda"), 77)
    assert float(acc[:100].abs().max()) == 0 and abs(float(acc[100:400].std()) - 2.0) < 0.3


def _build(task, patch):
    import yaml
    from msrflute_b200 import cli
    from msrflute_b200.core.config import FLUTEConfig
    import tempfile
    with open(os.path.join(ROOT, "experiments", task, "config.yaml")) as f:
        raw = yaml.safe_load(f)
    sc = raw["server_config"]
    sc["max_iteration"], sc["num_clients_per_iteration"] = 100, 3
    sc["val_freq"], sc["rec_freq"], sc["initial_val"], sc["initial_rec"] = 10 ** 9, 10 ** 9, False, False
    raw["model_config"]["compute_dtype"] = "fp32"
    patch(raw)
    config = FLUTEConfig.from_dict(raw)
    out = tempfile.mkdtemp(prefix="flute_dp_")
    mp = os.path.join(out, "models")
    os.makedirs(mp, exist_ok=True)
    config["data_path"], config["output_path"], config["model_path"], config["experiment_name"] = out, out, mp, "t"
    config["client_config"]["task"] = task
    config["server_config"]["task"] = task
    config.validate()
    return cli.build_job(config, task, out, mp)


def test_engine_local_dp_clip_matches_generic_path():
    """DGA + local DP in clip-only mode (eps < 0) is deterministic: the engine's fused gather must reproduce the
    generic per-client path (apply_local_dp on every client's flat pseudo-gradient)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def patch(raw):
        raw["strategy"] = "DGA"
        raw["server_config"]["aggregate_median"] = "mean"
        raw["dp_config"] = {"enable_local_dp": True, "enable_global_dp": False, "eps": -1.0, "max_grad": 0.05,
                            "max_weight": 1.0, "min_weight": 0.0, "delta": 1e-6, "global_sigma": 0.0}
        raw["client_config"]["data_config"]["train"]["batch_size"] = 4096      # one full-batch step per client

    server, worker, comm = _build("cv_lr_mnist", patch)
    assert worker.engine is not None and worker.engine.supports(worker.config)
    server.begin_training()
    from msrflute_b200.parallel.arena import module_arena
    w0 = module_arena(server.worker_trainer.model)[0].flat.clone()
    ids = [0, 3, 5]
    worker.set_weights(w0)
    worker.accumulator().zero_()
    outs_e = worker.engine.train_clients(ids, 0.05, 0, worker.weight_buffer(), worker.accumulator())
    acc_e = worker.accumulator().clone()
    worker.accumulator().zero_()
    eng, worker.engine = worker.engine, None
    outs_g = worker.train_clients(ids, (0.05, None, 0), fused=True)
    acc_g = worker.accumulator().clone()
    worker.engine = eng
    assert torch.allclose(acc_e, acc_g, rtol=2e-3, atol=1e-6), float((acc_e - acc_g).abs().max())
    # three clients, each clipped to max_grad
    assert float(acc_e.norm()) <= 3 * 0.05 * 1.001
    assert [o["pl"]["weight"] for o in outs_e] == [1.0, 1.0, 1.0]
    server.end_training()


@pytest.mark.parametrize("task", ["cv_lr_mnist", "cv_resnet_fedcifar100"])
def test_engine_fedprox_and_noisy_dp_rounds_run(task):
    """FedProx (closed-form proximal gradient) and DGA + local DP with Gaussian noise run on the engine, stay finite
    and the FedProx loss decreases."""
    def prox(raw):
        raw["strategy"] = "FedProx"
        raw["client_config"]["mu"] = 0.01
        if task == "cv_resnet_fedcifar100":
            raw["model_config"]["group_norm"] = 2

    server, worker, comm = _build(task, prox)
    assert worker.engine is not None and worker.engine.supports(worker.config)
    server.begin_training()
    losses = [server.run_rounds(1) for _ in range(5)]
    server.end_training()
    assert all(math.isfinite(v) for v in losses)
    assert worker.engine.prox_mult is not None
    assert min(losses[2:]) < losses[0], losses

    def dp(raw):
        raw["strategy"] = "DGA"
        raw["server_config"]["aggregate_median"] = "mean"
        raw["dp_config"] = {"enable_local_dp": True, "enable_global_dp": True, "eps": 100.0, "max_grad": 1.0,
                            "max_weight": 1.0, "min_weight": 0.0, "delta": 1e-6, "global_sigma": 0.001}
        if task == "cv_resnet_fedcifar100":
            raw["model_config"]["group_norm"] = 2

    server, worker, comm = _build(task, dp)
    assert worker.engine is not None and worker.engine.supports(worker.config)
    server.begin_training()
    losses = [server.run_rounds(1) for _ in range(3)]
    server.end_training()
    assert all(math.isfinite(v) for v in losses), losses


@pytest.mark.parametrize("q,bits", [(0.5, 8), (0.9, 4), (0.25, 2)])
def test_slot_quant_kernels_match_reference_quantizer(q, bits):
    """Radix-select statistics + fused binning == extensions/quantization applied to every client's materialised payload,
    including tensors whose dead filter taps (structural zeros) are not stored in the slot layout."""
    from msrflute_b200.extensions.quantization import quant
    from msrflute_b200.ops import arena_ops
    torch.manual_seed(3)
    S = 3
    # slot row: tensor A (1000 stored of 1000), tensor B (640 stored of 1500: 860 elided zeros), tensor C (33 of 33)
    offs, stored, total = [0, 1024, 1664], [1000, 640, 33], [1000, 1500, 33]
    Pc = 1728
    segs = torch.tensor([[o, n, t] for o, n, t in zip(offs, stored, total)], dtype=torch.int64, device="cuda")
    blk = torch.full((Pc // 32,), -1, dtype=torch.int16)
    for t, (o, n) in enumerate(zip(offs, stored)):
        blk[o // 32:(o + n + 31) // 32] = t
    blk = blk.cuda()
    wg = torch.zeros(Pc, device="cuda")
    W = torch.zeros(S, Pc, device="cuda")
    for o, n in zip(offs, stored):
        wg[o:o + n] = torch.randn(n, device="cuda")
        W[:, o:o + n] = wg[o:o + n] + torch.randn(S, n, device="cuda") * torch.rand(S, 1, device="cuda")
    coef = torch.tensor([1.0, 0.0, 2.5], device="cuda")
    params = arena_ops.slot_quant_stats(W, wg, segs, q, bits)
    acc = torch.zeros(Pc, device="cuda")
    arena_ops.slot_quant_gather(acc, W, wg, coef, params, blk, bits)
    want = torch.zeros(Pc, device="cuda")
    for s in range(S):
        c = float(coef[s])
        if c == 0:
            continue
        for o, n, t in zip(offs, stored, total):
            g = c * torch.cat([wg[o:o + n] - W[s, o:o + n], torch.zeros(t - n, device="cuda")])
            quant.quantize_tensor_(g, bits, q)
            want[o:o + n] += g[:n]
            assert float(g[n:].abs().max()) == 0.0 if t > n else True
    assert torch.allclose(acc, want, atol=1e-5, rtol=1e-4), float((acc - want).abs().max())


def test_engine_quantized_gather_matches_generic_path():
    """BASELINE config #4's algorithmic setting (DGA + gradient quantization) on the engine == the generic path."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def patch(raw):
        raw["strategy"] = "DGA"
        raw["server_config"]["aggregate_median"] = "mean"
        raw["client_config"]["quant_thresh"] = 0.5
        raw["client_config"]["quant_bits"] = 6
        raw["client_config"]["quant_anneal"] = 1.0
        raw["client_config"]["data_config"]["train"]["batch_size"] = 4096

    server, worker, comm = _build("cv_lr_mnist", patch)
    assert worker.engine is not None and worker.engine.supports(worker.config)
    server.begin_training()
    from msrflute_b200.parallel.arena import module_arena
    w0 = module_arena(server.worker_trainer.model)[0].flat.clone()
    ids = [1, 2, 9]
    worker.set_weights(w0)
    worker.accumulator().zero_()
    worker.engine.train_clients(ids, 0.05, 0, worker.weight_buffer(), worker.accumulator())
    acc_e = worker.accumulator().clone()
    worker.accumulator().zero_()
    eng, worker.engine = worker.engine, None
    worker.train_clients(ids, (0.05, None, 0), fused=True)
    acc_g = worker.accumulator().clone()
    worker.engine = eng
    server.end_training()
    # identical statistics up to float rounding of (w_global - w_local): allow a handful of boundary elements to land
    # in the neighbouring bin
    diff = (acc_e - acc_g).abs()
    assert float(diff.max()) < 2e-3 and float((diff > 1e-6).float().mean()) < 0.01, (float(diff.max()),)
    assert float(acc_e.norm()) > 0
