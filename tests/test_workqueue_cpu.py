"""Job-wide atomic counter (``Communicator.fetch_add``) behind the ``work_queue`` dispatch policy, and the eager fallback
of the CUDA-graphed client step off the GPU."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fetch_add_hands_out_every_index_exactly_once_across_ranks(tmp_path):
    script = tmp_path / "pull.py"
    script.write_text(textwrap.dedent('''
        import json, os, sys
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from msrflute_b200.parallel.comm import CollectiveComm
        from msrflute_b200.core.federated import WorkQueue
        dist.init_process_group("gloo")
        comm = CollectiveComm()
        got = [comm.fetch_add("t/a", 1) for _ in range(50)]
        wq = WorkQueue(comm, "t/q", list(range(100, 137)))
        taken = list(wq)
        comm.barrier()
        with open(os.path.join(%r, "r%%d.json" %% comm.rank), "w") as f:
            json.dump({"got": got, "taken": taken}, f)
        dist.destroy_process_group()
    ''' % (ROOT, str(tmp_path))))
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr",
                        "127.0.0.1", "--master-port", "29817", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    res = [json.load(open(tmp_path / ("r%d.json" % k))) for k in range(3)]
    allgot = sorted(v for x in res for v in x["got"])
    assert allgot == list(range(150))                              # atomic: no index handed out twice, none skipped
    alltaken = sorted(v for x in res for v in x["taken"])
    assert alltaken == list(range(100, 137))                       # every queued client trained exactly once


def test_local_comm_counter_and_graphed_step_eager_fallback_on_cpu():
    import torch
    from msrflute_b200.parallel.comm import LocalComm
    from msrflute_b200.core.federated import WorkQueue
    c = LocalComm(device=torch.device("cpu"))
    assert [c.fetch_add("k") for _ in range(3)] == [0, 1, 2] and c.fetch_add("other", 5) == 0
    assert list(WorkQueue(c, "q", [7, 8, 9])) == [7, 8, 9]
    from msrflute_b200.core.graphed import GraphedTrainStep

    class _T:                                     # minimal trainer face: not capturable on CPU -> eager step
        use_arena = False
        optimizer = None
        ss_scheduler = None
        model = torch.nn.Linear(2, 2)
        calls = 0

        def _train_step(self, batch):
            self.calls += 1
            return torch.zeros(())

    t = _T()
    gs = GraphedTrainStep(t)
    gs({"x": torch.zeros(4, 2)})
    gs({"x": torch.zeros(4, 2)})
    assert t.calls == 2 and gs.fallbacks == 2 and gs.captures == 0
