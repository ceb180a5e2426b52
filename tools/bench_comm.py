"""Roofline check of the fused communication kernels (run under torch.distributed.run, one rank per GPU):

  p2p_broadcast                 server stores the 11.7 M-float arena into every peer      egress  (W-1)*P*4 B
  fused_reduce_update_bcast     P2P loads of every accumulator + SGD + P2P stores         ingress (W-1)*P*4 B + egress (W-1)*P*4 B
  NCCL broadcast + reduce       the baseline transport for the same bytes

Timed on the server GPU with CUDA events between device barriers; prints one JSON line from rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from msrflute_b200.ops import _ext, arena_ops
from msrflute_b200.parallel.comm import init_distributed, make_communicator

NVLINK_GBS = float(os.environ.get("NVLINK_GBS", "770"))          # measured per-direction NVLink-5 copy bandwidth


def timed(fn, sync, iters=10, warmup=3):
    ts = []
    for i in range(warmup + iters):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    rank, world = init_distributed("nccl")
    comm = make_communicator("symm")
    ext = _ext.load(required=True)
    P = 11_699_136
    w = comm.alloc_flat(P, name="w_global")
    acc = comm.alloc_flat(P, name="acc")
    acc.fill_(float(rank + 1))
    w.normal_()
    peers_w, peers_acc = comm.peer_weight_buffers(w), comm.peer_accumulators(acc)
    opt = arena_ops.ServerOptState("sgd", P, w.device, lr=1.0)
    wsum = torch.tensor(float(world), device=w.device)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    out = {"world": world, "P": P, "bytes_per_peer": P * 4}
    if rank == 0:
        ms = timed(lambda: ext.p2p_broadcast(w, peers_w), barrier)
        out["p2p_broadcast_ms"] = ms
        out["p2p_broadcast_egress_GBs"] = (world - 1) * P * 4 / ms / 1e6
        ms = timed(lambda: arena_ops.server_update(w, peers_acc, wsum, opt, bcast=peers_w, zero_accs=False), barrier)
        out["fused_update_ms"] = ms
        out["fused_update_ingress_GBs"] = (world - 1) * P * 4 / ms / 1e6
        out["fused_update_frac_of_nvlink"] = out["fused_update_ingress_GBs"] / NVLINK_GBS
        out["p2p_broadcast_frac_of_nvlink"] = out["p2p_broadcast_egress_GBs"] / NVLINK_GBS
    else:
        for _ in range(2 * 13):
            barrier()
    # NCCL baseline: broadcast + reduce of the same arena (every rank participates)
    ms = timed(lambda: (dist.broadcast(w, src=0), dist.reduce(acc, dst=0)), barrier)
    if rank == 0:
        out["nccl_broadcast_plus_reduce_ms"] = ms
        print(json.dumps(out))
    barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
