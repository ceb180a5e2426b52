"""Multi-GPU checks launched under torch.distributed.run by tests/test_multigpu.py.

    --mode fl      : R federated rounds of the flagship config on a small synthetic population; rank 0 prints a JSON line
                     with the loss trajectory and a checksum of the final global weights (compare transports).
    --mode stress  : 1000 rounds of the raw sharded transport (SymmComm.sharded_round) on rank-dependent accumulators;
                     every rank checks its weight buffer against the closed form every round (flag / barrier stress).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def fl(args):
    import logging
    import random
    logging.getLogger().setLevel(logging.WARNING)
    os.chdir(ROOT)
    import bench
    from msrflute_b200.core.federated import Server
    # LR-MNIST: deterministic kernels only (no atomics in the client step), so transports can be compared tightly; the
    # flagship's tf32 / atomic-order noise is amplified by its last-stage GroupNorm and would mask transport bugs
    if args.task == "cv_resnet_fedcifar100":
        job = bench.build_flagship(n_clients_per_round=args.clients, users=40, norm="gn", comm=args.comm)
    else:
        job = bench.build_job(args.task, args.clients, comm=args.comm, compute_dtype="fp32")
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        job.worker.run()
        job.comm.close()
        return
    random.seed(7)
    torch.manual_seed(7)
    srv = job.server
    srv.begin_training()
    losses = [srv.run_rounds(1) for _ in range(args.rounds)]
    Server.sync_nodes()
    from msrflute_b200.parallel.arena import module_arena
    w = module_arena(srv.worker_trainer.model)[0].flat.double()
    out = {"comm": job.comm.kind, "sharded_mode": getattr(job.comm, "last_sharded_mode", None), "losses": losses,
           "w_norm": float(w.norm()), "w_sum": float(w.sum()), "w_probe": [float(v) for v in w[100:108]]}
    srv.end_training()
    print("MGCHECK " + json.dumps(out), flush=True)
    job.comm.close()


def stress(args):
    import torch.distributed as dist
    from msrflute_b200.parallel.comm import init_distributed, make_communicator
    init_distributed("nccl")
    comm = make_communicator("symm")
    assert comm.kind == "symm", comm.kind
    rank, N = comm.rank, comm.size
    P = 1 << 20
    w = comm.alloc_flat(P, torch.float32, name="w")
    acc = comm.alloc_flat(P, torch.float32, name="acc")
    w.fill_(1.0)
    comm.device_barrier()
    expect = torch.ones(P, device=w.device, dtype=torch.float64)
    idx = torch.arange(P, device=w.device, dtype=torch.float64)
    bad = 0
    opt = {"code": 0, "step": 1, "lr": 0.5}
    for r in range(args.rounds):
        acc.copy_(((idx % 97) * (rank + 1) * 1e-3 + r * 1e-4).float())
        wsum = torch.tensor([float(rank + 1)], device=w.device)
        comm.sharded_round(w, acc, wsum, dict(opt, step=r + 1))
        tot_w = N * (N + 1) / 2.0
        g = ((idx % 97) * 1e-3 * tot_w + N * r * 1e-4) / tot_w
        expect -= 0.5 * g
        if r % 50 == 49 or r == args.rounds - 1:
            err = float((w.double() - expect).abs().max())
            if err > 1e-2 or float(acc.abs().max()) != 0.0:
                bad += 1
    t = torch.tensor([bad], device=w.device)
    dist.all_reduce(t)
    if rank == 0:
        print("MGCHECK " + json.dumps({"stress_rounds": args.rounds, "bad_checks": int(t.item()),
                                       "sharded_mode": getattr(comm, "last_sharded_mode", None)}), flush=True)
    comm.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="fl")
    ap.add_argument("--comm", default="auto")
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--clients", type=int, default=6)
    ap.add_argument("--task", default="cv_lr_mnist")
    a = ap.parse_args()
    (fl if a.mode == "fl" else stress)(a)
