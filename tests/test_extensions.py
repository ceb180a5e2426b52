import math

import numpy as np
import pytest
import torch

from msrflute_b200.extensions import privacy
from msrflute_b200.extensions.privacy import analysis
from msrflute_b200.extensions.quantization import quant as Q
from msrflute_b200.ops import quant_ops
from msrflute_b200.parallel.arena import adopt_module


class _T:
    def __init__(self, model):
        self.model = model


def _model_with_grads(scale=1.0, seed=0):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(40, 30), torch.nn.Linear(30, 20))
    adopt_module(m)
    for p in m.parameters():
        p.grad.copy_(torch.randn_like(p) * scale)
    return m


def _flat_grad(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def test_local_dp_clip_only_when_eps_negative():
    m = _model_with_grads(scale=5.0)
    n0 = _flat_grad(m).norm().item()
    w = privacy.apply_local_dp(_T(m), 2.0, {"eps": -1.0, "max_grad": 1.0}, add_weight_noise=False)
    assert w == 2.0 and n0 > 1.0
    assert math.isclose(_flat_grad(m).norm().item(), 1.0, rel_tol=1e-4)
    m2 = _model_with_grads(scale=1e-3)
    before = _flat_grad(m2).clone()
    privacy.apply_local_dp(_T(m2), 1.0, {"eps": -1.0, "max_grad": 1.0}, add_weight_noise=False)
    assert torch.allclose(_flat_grad(m2), before)             # already inside the ball: untouched


def test_local_dp_noise_scale_matches_gaussian_mechanism():
    torch.manual_seed(0)
    m = torch.nn.Linear(400, 250)
    adopt_module(m)
    for p in m.parameters():
        p.grad.copy_(torch.randn_like(p))
    direction = _flat_grad(m) / _flat_grad(m).norm()
    cfg = {"eps": 4.0, "delta": 1e-6, "max_grad": 2.0, "max_weight": 10.0, "min_weight": 0.01, "weight_scaler": 1.0}
    w = privacy.apply_local_dp(_T(m), 3.0, cfg, add_weight_noise=True)
    sens = math.sqrt(2.0 ** 2 + 10.0 ** 2)
    sigma = privacy.compute_LDP_noise_std(4.0, sens, 1e-6)
    resid = _flat_grad(m) - 2.0 * direction                   # signal was normalised to max_grad
    assert abs(resid.std().item() / sigma - 1.0) < 0.02
    assert 0.01 <= w <= 10.0
    assert privacy.apply_local_dp(_T(m), 3.0, cfg, add_weight_noise=False) == 3.0


def test_global_dp_noise_scale_and_norm_logging():
    m = _model_with_grads(scale=0.0)
    cfg = {"dp_config": {"enable_local_dp": True, "enable_global_dp": True, "global_sigma": 0.5, "max_grad": 4.0}}
    logged = {}
    scale = privacy.apply_global_dp(cfg, m, num_clients_curr_iter=8, metric_logger=lambda k, v: logged.__setitem__(k, v))
    assert math.isclose(scale, 0.5 * 4.0 / 8)
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert abs(g.std().item() / scale - 1.0) < 0.05 and logged["Gradient Norm"] == 0.0
    assert privacy.apply_global_dp({"dp_config": {"enable_local_dp": False}}, m, 8) is None


def test_rdp_accountant_closed_forms_and_monotonicity():
    # q = 1: plain Gaussian mechanism, RDP(alpha) = alpha / (2 sigma^2)
    assert math.isclose(analysis.compute_rdp(1.0, 2.0, 10, 8.0), 10 * 8.0 / (2 * 4.0))
    assert analysis.compute_rdp(0.0, 2.0, 10, 8.0) == 0
    orders = privacy.ORDERS
    eps = []
    for steps in (1, 10, 100):
        rdp = analysis.compute_rdp(0.01, 1.1, steps, orders)
        e, a = analysis.get_privacy_spent(orders, rdp, 1e-5)
        eps.append(e)
        assert a in orders
    assert eps[0] < eps[1] < eps[2]
    # published sanity point (TF-privacy MNIST tutorial): q=256/60000, sigma=1.1, 60 epochs, delta=1e-5 -> eps = 3.01
    steps = int(60 * 60000 / 256)
    e, _ = analysis.get_privacy_spent(orders, analysis.compute_rdp(256 / 60000, 1.1, steps, orders), 1e-5)
    assert abs(e - 3.01) < 0.02, e
    # integer and fractional orders agree at the boundary
    a = analysis._compute_rdp(0.05, 1.3, 4.0)
    b = analysis._compute_rdp(0.05, 1.3, 4.0000001)
    assert abs(a - b) < 1e-4


def test_update_privacy_accountant_logs_and_returns_eps():
    cfg = {"dp_config": {"enable_local_dp": True, "enable_global_dp": True, "global_sigma": 1.0, "max_grad": 1.0,
                         "delta": 1e-6}}
    logged = {}
    e1 = privacy.update_privacy_accountant(cfg, 1000, 0, 10, metric_logger=lambda k, v: logged.__setitem__(k, v))
    e2 = privacy.update_privacy_accountant(cfg, 1000, 49, 10, metric_logger=lambda k, v: logged.__setitem__(k, v))
    assert 0 < e1 < e2 and logged["dp_global_T"] == 50 and logged["dp_global_B"] == 10
    assert privacy.update_privacy_accountant({"dp_config": {"enable_local_dp": False}}, 10, 0, 1) is None


def _reference_quant(g, bits, thr):
    """The reference algorithm verbatim in spirit: linspace levels + bucketize(g − w/2) + threshold (quant.py:53-100)."""
    lo, hi = g.min(), g.max()
    th = torch.quantile(g.abs(), thr)
    labels = torch.linspace(lo, hi, 2 ** bits)
    width = labels[1] - labels[0]
    binned = labels[torch.bucketize(g - 0.5 * width, labels, right=False).clamp(max=2 ** bits - 1)]
    return torch.where(g.abs() > th, binned, torch.tensor(0.0))


@pytest.mark.parametrize("bits,thr", [(4, 0.5), (8, 0.9), (10, 0.0)])
def test_quantization_matches_reference_algorithm(bits, thr):
    torch.manual_seed(0)
    g = torch.randn(5000)
    want = _reference_quant(g.clone(), bits, thr)
    got = Q.quantize_tensor_(g.clone(), bits, thr)
    width = (g.max() - g.min()) / (2 ** bits - 1)
    mism = (got - want).abs() > 1e-5
    # ties on a bin edge may round either way in float arithmetic; they must be rare and one bin wide
    assert mism.float().mean() < 2e-3
    assert ((got - want).abs()[mism] <= width * 1.001).all()
    assert (got == 0).float().mean() >= thr - 1e-3
    assert len(torch.unique(got)) <= 2 ** bits + 1


def test_quant_model_on_arena_and_pack_roundtrip():
    m = _model_with_grads()
    ref = [Q.quantize_tensor_(p.grad.clone(), 6, 0.7) for p in m.parameters()]
    Q.quant_model(m, quant_bits=6, quant_threshold=0.7)
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p.grad, r, atol=1e-6)
    Q.quant_model(m, quant_bits=6, quant_threshold=None)          # no-op
    flat = torch.randn(4096)
    segs = [(0, 1000), (1024, 3000)]
    codes, bitmap, stats = quant_ops.pack_segments(flat, segs, 6, 0.7)
    sim = quant_ops.quantize_segments_(flat.clone(), segs, 6, 0.7)
    out = torch.zeros(4096)
    quant_ops.unpack_add_(out, codes, bitmap, stats, segs, 6, alpha=1.0)
    for o, n in segs:
        assert torch.allclose(out[o:o + n], sim[o:o + n], atol=1e-5)
    assert bitmap.numel() * 8 >= 4096 and codes.dtype == torch.uint8      # 6-bit codes + 1 bit/elem on the wire


def test_dp_kmeans_recovers_well_separated_clusters():
    from msrflute_b200.extensions.privacy.dp_kmeans import DPKMeans, sample, sphere_packing_initialization
    np.random.seed(0)
    centers = np.array([[0.8, 0.0], [-0.8, 0.0], [0.0, 0.8]])
    X = np.concatenate([c + 0.05 * np.random.randn(400, 2) for c in centers])
    km = DPKMeans(n_dim=2, eps=50.0, max_cluster_l2=1.0, n_clusters=3, max_iter=15, device="cpu")
    labels = km.fit_predict(X)
    assert km.eps[0] == 50.0 * km.n_iter_
    found = km.cluster_centers_
    for c in centers:
        assert np.min(np.linalg.norm(found - c, axis=1)) < 0.2
    assert len(set(labels[:400])) == 1
    pts = sample(5, 2.0, 100)
    assert (np.linalg.norm(pts, axis=1) <= 2.0 + 1e-9).all()
    init, a = sphere_packing_initialization(4, 3, 0.2, 1.0, 50)
    d = np.linalg.norm(init[:, None] - init[None], axis=-1) + np.eye(4) * 10
    assert d.min() >= 2 * a - 1e-9


def test_rl_agent_forward_train_save_load(tmp_path):
    from msrflute_b200.extensions.RL import RL
    cfg = {"num_clients_per_iteration": 4,
           "RL": {"RL_path": str(tmp_path), "initial_epsilon": 0.0, "final_epsilon": 0.0, "epsilon_gamma": 0.9,
                  "max_replay_memory_size": 10, "minibatch_size": 4, "network_params": "16,32,4",
                  "optimizer_config": {"type": "adam", "lr": 0.01},
                  "annealing_config": {"type": "step_lr", "step_interval": "epoch", "gamma": 1.0, "step_size": 100}}}
    rl = RL(config=cfg)
    state = np.random.rand(16)
    a = rl.forward(state)
    assert a.shape == (4,)
    rl.set_weights(np.exp(a.detach().cpu().numpy()))
    rl.set_losses((1.0, 0.5))
    for _ in range(6):
        rl.train((state, rl.rl_weights, [1.0]))
    assert rl.runningLoss > 0
    rl.save(3)
    rl2 = RL(config=cfg)
    assert rl2.cur_iter_no == 4
    for p, q in zip(rl.model.parameters(), rl2.model.parameters()):
        assert torch.equal(p.cpu(), q.cpu())


def test_lr_scheduler_factory_accepts_schema_mandated_keys():
    """gamma / step_size are required by the server annealing schema; every scheduler type must tolerate them."""
    import torch
    from msrflute_b200.utils import make_lr_scheduler
    p = [torch.nn.Parameter(torch.zeros(1))]
    base = {"step_interval": "epoch", "gamma": 0.5, "step_size": 2}
    s = make_lr_scheduler(dict(base, type="step_lr"), torch.optim.SGD(p, lr=1.0))
    assert isinstance(s, torch.optim.lr_scheduler.StepLR)
    s = make_lr_scheduler(dict(base, type="multi_step_lr", milestones=[1, 3]), torch.optim.SGD(p, lr=1.0))
    assert isinstance(s, torch.optim.lr_scheduler.MultiStepLR) and list(s.milestones) == [1, 3]
    s = make_lr_scheduler(dict(base, type="rampup-keep-expdecay-keep", peak_lr=0.05, floor_lr=0.001, sr=1, si=2, sf=3),
                          torch.optim.SGD(p, lr=1.0))
    assert abs(s.lr_at(1) - 0.05) < 1e-9 and abs(s.lr_at(10) - 0.001) < 1e-9
    opt = torch.optim.SGD(p, lr=1.0)
    s = make_lr_scheduler(dict(base, type="val_loss"), opt)
    assert isinstance(s, torch.optim.lr_scheduler.ReduceLROnPlateau) and s.factor == 0.5 and s.patience == 2
    for loss in (1.0, 1.0, 1.0, 1.0):
        s.step(loss)
    assert opt.param_groups[0]["lr"] == 0.5


@pytest.mark.parametrize("bits", [4, 8, 10])
def test_quantised_payload_wire_format_round_trips(bits):
    """ops.quant_ops.wire_encode / wire_decode: level codes + keep bitmap + (lo, width) table reproduce the simulated-
    quantization values, at a fraction of the fp32 bytes."""
    import torch
    from msrflute_b200.ops import quant_ops
    torch.manual_seed(bits)
    sizes = [1000, 37, 4096]
    offs = [0, 1000, 1037]
    flat = torch.randn(sum(sizes)) * torch.repeat_interleave(torch.tensor([1.0, 0.01, 30.0]), torch.tensor(sizes))
    segs = list(zip(offs, sizes))
    q, stats = quant_ops.quantize_segments_(flat.clone(), segs, bits, 0.6, return_stats=True)
    codes, bitmap, table = quant_ops.wire_encode(q, sizes, stats[:, :2], bits)
    back = quant_ops.wire_decode(codes, bitmap, table, sizes)
    assert torch.allclose(back, q, rtol=1e-5, atol=1e-6 * float(q.abs().max()))
    assert float((back == 0).float().mean()) > 0.5                      # the thresholded 60 % stay exactly zero
    wire = codes.numel() * codes.element_size() + bitmap.numel() + table.numel() * 4
    assert wire < (0.30 if bits <= 8 else 0.55) * q.numel() * 4
