import copy
import math
import os

import numpy as np
import pytest
import torch
import yaml

from msrflute_b200.core.config import FLUTEConfig
from msrflute_b200.core.strategies import select_strategy, FedAvg, DGA, FedLabels
from msrflute_b200.core.strategies.utils import filter_weight, filter_weight_tensor
from msrflute_b200.core.trainer import ModelUpdater
from msrflute_b200.parallel.arena import adopt_module, module_arena
from msrflute_b200.utils import make_optimizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config(**over):
    with open(os.path.join(ROOT, "experiments", "cv_lr_mnist", "config.yaml")) as f:
        raw = yaml.safe_load(f)
    cfg = FLUTEConfig.from_dict(raw)
    cfg["model_path"] = over.pop("model_path", "/tmp")
    for k, v in over.items():
        node = cfg
        ks = k.split(".")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    return cfg


class _FakeTrainer:
    def __init__(self, model, num_samples, train_loss, stats=None):
        self.model, self.num_samples, self.train_loss = model, num_samples, train_loss
        self.sufficient_stats = stats or {"var": 0.1, "mean": 0.2, "mag": 0.3}
        self.algo_computation = None

    def reset_gradient_power(self): pass
    def estimate_sufficient_stats(self): pass


def _model(seed=0):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
    adopt_module(m)
    return m


def _updater(model, lr=1.0):
    opt = make_optimizer({"type": "sgd", "lr": lr}, model)
    return ModelUpdater(model, opt, None, None, None, None, {"type": "step_lr", "step_interval": "epoch", "gamma": 1.0, "step_size": 100})


def test_select_strategy_mapping():
    assert select_strategy("DGA") is DGA and select_strategy("FedAvg") is FedAvg
    assert select_strategy("FedProx") is FedAvg and select_strategy("fedlabels") is FedLabels
    with pytest.raises(ValueError):
        select_strategy("nope")


def test_filter_weight_clamps():
    assert filter_weight(float("nan")) == 0.0 and filter_weight(float("inf")) == 0.0
    assert filter_weight(250.0) == 100 and filter_weight(3.5) == 3.5
    t = filter_weight_tensor(torch.tensor([float("nan"), float("inf"), 250.0, 3.5]))
    assert t.tolist() == [0.0, 0.0, 100.0, 3.5]


@pytest.mark.parametrize("fast", [False, True])
def test_fedavg_equals_closed_form(fast):
    """w_new = w − lr_server · Σ_k n_k (w − w_k) / Σ_k n_k  (server SGD lr 1 ⇒ sample-weighted model average)."""
    cfg = _config(**{"server_config.fast_aggregation": fast})
    server_model = _model(0)
    w0 = [p.detach().clone() for p in server_model.parameters()]
    updater = _updater(server_model)
    srv = FedAvg("server", cfg, "/tmp")
    cli = FedAvg("client", cfg)
    ns = [10, 30, 60]
    locals_ = []
    for k, n in enumerate(ns):
        cm = _model(0)
        with torch.no_grad():
            for p in cm.parameters():
                p.add_(torch.randn_like(p) * 0.1 * (k + 1))
        locals_.append([p.detach().clone() for p in cm.parameters()])
        for p, g in zip(cm.parameters(), w0):
            p.grad.copy_(g - p.data)
        payload = cli.generate_client_payload(_FakeTrainer(cm, n, 1.0))
        assert payload["weight"] == n
        assert srv.process_individual_payload(updater, payload)
    assert not srv.process_individual_payload(updater, {"weight": 0.0, "gradients": []})    # dropped client
    srv.combine_payloads(updater, 0, 3, 100, None)
    for i, p in enumerate(server_model.parameters()):
        expect = sum(n * l[i] for n, l in zip(ns, locals_)) / sum(ns)
        assert torch.allclose(p.data, expect, atol=1e-5)


def test_dga_softmax_weight_and_mean_mode():
    cfg = _config(**{"strategy": "DGA", "server_config.aggregate_median": "softmax", "server_config.softmax_beta": 2.0})
    cli = DGA("client", cfg)
    m = _model(1)
    for p in m.parameters():
        p.grad.fill_(1.0)
    tr = _FakeTrainer(m, 20, 10.0)
    payload = cli.generate_client_payload(tr)
    assert math.isclose(payload["weight"], math.exp(-2.0 * 10.0 / 20), rel_tol=1e-6)
    assert torch.allclose(payload["flat"][:24], torch.full((24,), payload["weight"]))
    for kind, key in (("mag_var_loss", "var"), ("mag_mean_loss", "mean"), ("other", "mag")):
        cfg2 = _config(**{"strategy": "DGA", "server_config.aggregate_median": "softmax", "server_config.softmax_beta": 1.0,
                          "server_config.weight_train_loss": kind})
        w = DGA("client", cfg2).client_weight(tr)
        assert math.isclose(w, math.exp(-tr.sufficient_stats[key]), rel_tol=1e-6)
    cfg3 = _config(**{"strategy": "DGA", "server_config.aggregate_median": "mean"})
    assert DGA("client", cfg3).generate_client_payload(_FakeTrainer(_model(2), 5, 1.0))["weight"] == 1.0


def test_dga_stale_gradients_are_deferred_not_lost():
    cfg = _config(**{"strategy": "DGA", "server_config.aggregate_median": "mean", "server_config.stale_prob": 1.0})
    model = _model(0)
    w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    updater = _updater(model)
    srv = DGA("server", cfg, "/tmp")
    assert srv.needs_individual_payloads
    g = [torch.ones_like(p) for p in model.parameters()]
    logged = {}
    srv.process_individual_payload(updater, {"weight": 1.0, "gradients": [x.clone() for x in g]})
    np.random.seed(0)
    srv.combine_payloads(updater, 0, 1, 10, None, logger=lambda k, v: logged.__setitem__(k, v))
    w1 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(w1, w0)                         # everything went stale: no update this round
    assert len(srv.client_parameters_stack_stale) == 1 and srv.weight_sum_stale == 1.0
    srv.stale_prob = 0.0
    srv.process_individual_payload(updater, {"weight": 1.0, "gradients": [x.clone() for x in g]})
    srv.combine_payloads(updater, 1, 1, 10, None, logger=lambda k, v: logged.__setitem__(k, v))
    w2 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(w2, w0 - 1.0, atol=1e-6)        # (stale + fresh) / (1 + 1) = 1 ⇒ w −= 1
    assert logged["Stale Gradients Ratio"] == 1.0


def test_fedlabels_average_of_sup_and_weighted_unsup():
    cfg = _config(**{"strategy": "FedLabels"})
    model = _model(0)
    updater = _updater(model)
    srv = FedLabels("server", cfg, "/tmp")
    keys = list(model.state_dict().keys())
    sups = [[torch.full_like(model.state_dict()[k], float(v)) for k in keys] for v in (1.0, 3.0)]
    unsups = [[torch.full_like(model.state_dict()[k], float(v)) for k in keys] for v in (10.0, 20.0)]
    for w, s, u in zip((1, 3), sups, unsups):
        srv.process_individual_payload(updater, {"weight": w, "gradients": s + u})
    srv.combine_payloads(updater, 0, 2, 10, None)
    sup_avg, unsup_avg = (1.0 + 3.0) / 2, (1 * 10.0 + 3 * 20.0) / 4
    for p in model.parameters():
        assert torch.allclose(p.data, torch.full_like(p, (sup_avg + unsup_avg) / 2))
    assert set(srv.tmp_sup.keys()) == set(keys)


def test_dispatch_policies_lpt_round_robin_dynamic():
    from msrflute_b200.core import federated as F
    items, costs = list(range(8)), [8, 7, 6, 5, 4, 3, 2, 1]
    rr = F.assign_clients(items, costs, [0, 1], "round_robin")
    assert rr == {0: [0, 2, 4, 6], 1: [1, 3, 5, 7]}
    lpt = F.assign_clients(items, costs, [0, 1], "static_lpt")
    assert abs(sum(costs[i] for i in lpt[0]) - sum(costs[i] for i in lpt[1])) <= 1
    # dynamic: a worker measured 3x slower gets ~1/4 of the work
    F._Runtime.options.pop("_worker_speeds", None)
    F.update_worker_speeds({0: [0, 1], 1: [2, 3]}, {0: 10, 1: 10, 2: 10, 3: 10}, {0: 1.0, 1: 3.0})
    dyn = F.assign_clients(items, costs, [0, 1], "dynamic", speeds=F._Runtime.options["_worker_speeds"])
    load = {w: sum(costs[i] for i in v) for w, v in dyn.items()}
    assert load[0] > 2.3 * load[1] and sorted(dyn[0] + dyn[1]) == items
    F._Runtime.options.pop("_worker_speeds", None)
