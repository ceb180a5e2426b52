"""core/graphed.py: the CUDA-graph-captured generic client step reproduces the eager step."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from test_engine_dp_prox_gpu import _build        # noqa: E402


def test_graphed_generic_step_matches_eager():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def patch(raw):
        raw["client_config"]["data_config"]["train"]["batch_size"] = 16
        raw["client_config"]["optimizer_config"]["momentum"] = 0.9

    server, worker, comm = _build("cv_lr_mnist", patch)
    server.begin_training()
    from msrflute_b200.parallel.arena import module_arena
    from msrflute_b200.core.client import ClientContext
    w0 = module_arena(server.worker_trainer.model)[0].flat.clone()
    ids = [1, 2, 9, 4]
    eng, worker.engine = worker.engine, None
    accs = {}
    for mode in ("0", "1"):
        os.environ["FLUTE_GRAPHED_STEP"] = mode
        import random
        import numpy as np
        torch.manual_seed(11); np.random.seed(11); random.seed(11)      # same mini-batch order in both runs
        worker.set_weights(w0)
        worker.accumulator().zero_()
        worker.train_clients(ids, (0.05, None, 0), fused=True)
        accs[mode] = worker.accumulator().clone()
    del os.environ["FLUTE_GRAPHED_STEP"]
    gs = ClientContext.of(worker.model).graphed
    worker.engine = eng
    server.end_training()
    assert gs is not None and gs.captures >= 1 and gs.replays > 4 * gs.captures, (gs.captures, gs.replays, gs.fallbacks)
    assert float(accs["0"].norm()) > 0
    assert torch.allclose(accs["0"], accs["1"], rtol=1e-5, atol=1e-7), float((accs["0"] - accs["1"]).abs().max())
