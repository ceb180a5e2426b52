"""Rank-to-rank transports for the two communication-bound paths of an FL round.

The reference moves the model with ``1 + 3·n_tensors`` point-to-point messages per worker per direction, every
tensor staged through the CPU on both sides (``core/federated.py:112-124,330-334``).  Here the
model is ONE flat arena, and a round needs exactly two data-plane operations:

``broadcast_weights``   server arena → every rank's arena
``reduce_accumulators`` Σ_ranks (weighted pseudo-gradient accumulators) → server

Three interchangeable implementations:

* :class:`LocalComm`       world_size == 1 (everything aliases, no-ops).
* :class:`CollectiveComm`  ``torch.distributed`` broadcast/reduce on the flat buffers — gloo (CPU plumbing,
                           BASELINE config #1) or NCCL (the *baseline to beat* on GPUs).
* :class:`SymmComm`        (``parallel/symm.py``) hand-written kernels over NVLink peer memory: the server's
                           reduce+optimizer kernel loads peers' accumulators with P2P ``ld.global`` and
                           stores the new weights straight into every peer's arena — no NCCL on the path.

Control-plane messages (commands, client assignments, metrics dicts) are small python objects sent over a
CPU-side gloo group so they never synchronise the GPU stream.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional

import torch
import torch.distributed as dist

from ..utils import env_local_rank, env_rank, env_world_size, print_rank


class Communicator:
    kind = "base"

    def __init__(self):
        self.rank, self.size = env_rank() if dist.is_initialized() else 0, 1
        self.device = torch.device("cpu")

    # ---- control plane -------------------------------------------------
    def bcast_object(self, obj: Any, src: int = 0) -> Any:
        return obj

    def gather_objects(self, obj: Any) -> List[Any]:
        """all-gather of python objects (every rank gets the list)."""
        return [obj]

    def barrier(self):
        pass

    # ---- data plane ----------------------------------------------------
    def alloc_flat(self, numel: int, dtype=torch.float32, name: str = "") -> torch.Tensor:
        return torch.zeros(numel, dtype=dtype, device=self.device)

    def broadcast_weights(self, w: torch.Tensor, src: int = 0):
        pass

    def reduce_accumulators(self, acc: torch.Tensor, dst: int = 0):
        pass

    def gather_flat(self, t: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
        """Gather equally-sized flat tensors on ``dst`` (individual-payload mode)."""
        return [t]

    def all_reduce_(self, t: torch.Tensor):
        return t

    def close(self):
        pass


class LocalComm(Communicator):
    kind = "local"

    def __init__(self, device=None):
        self.rank, self.size = 0, 1
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else torch.device("cpu"))
        self._counters = {}

    def fetch_add(self, key: str, n: int = 1) -> int:
        """Value of the job-wide counter ``key`` before adding ``n`` (the work-queue primitive; trivially local here)."""
        v = self._counters.get(key, 0)
        self._counters[key] = v + n
        return v


class CollectiveComm(Communicator):
    """torch.distributed collectives on flat buffers; control plane on a gloo side-group."""
    kind = "collective"

    def __init__(self, device=None):
        assert dist.is_initialized()
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.backend = dist.get_backend()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        self.device = device
        # CPU-side control group: object collectives never touch the GPU stream
        import datetime
        timeout = datetime.timedelta(seconds=float(os.environ.get("FLUTE_COMM_TIMEOUT_S", "1800")))
        self.ctrl_group = dist.new_group(backend="gloo", timeout=timeout) if self.backend != "gloo" else dist.group.WORLD

    def fetch_add(self, key: str, n: int = 1) -> int:
        """Atomic fetch-and-add on a job-wide counter kept by the rendezvous store (one TCP round trip, no GPU work):
        the ``work_queue`` dispatch policy pulls client indices from it, so an idle rank takes the next client the
        moment it finishes one — what the reference does with per-client acknowledgements (``federated.py:344-373``)."""
        from torch.distributed import distributed_c10d as c10d
        store = c10d._get_default_store()
        return int(store.add("flute/" + key, int(n))) - int(n)

    def bcast_object(self, obj, src=0):
        box = [obj]
        dist.broadcast_object_list(box, src=src, group=self.ctrl_group)
        return box[0]

    def gather_objects(self, obj):
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.ctrl_group)
        return out

    def barrier(self):
        dist.barrier(group=self.ctrl_group)

    def _staged(self, t):
        """gloo cannot move CUDA tensors; stage through the host in that (test-only) combination."""
        return t.is_cuda and self.backend == "gloo"

    def broadcast_weights(self, w, src=0):
        if self._staged(w):
            h = w.cpu()
            dist.broadcast(h, src=src)
            w.copy_(h)
        else:
            dist.broadcast(w, src=src)

    def reduce_accumulators(self, acc, dst=0):
        if self._staged(acc):
            h = acc.cpu()
            dist.reduce(h, dst=dst, op=dist.ReduceOp.SUM)
            if self.rank == dst:
                acc.copy_(h)
        else:
            dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)

    def all_reduce_(self, t):
        if self._staged(t):
            h = t.cpu()
            dist.all_reduce(h)
            t.copy_(h)
        else:
            dist.all_reduce(t)
        return t

    def gather_flat(self, t, dst=0):
        staged = self._staged(t)
        src_t = t.cpu() if staged else t
        bufs = [torch.empty_like(src_t) for _ in range(self.size)] if self.rank == dst else None
        dist.gather(src_t, gather_list=bufs, dst=dst)
        if bufs is None:
            return None
        return [b.to(t.device) for b in bufs] if staged else bufs


def init_distributed(backend: Optional[str] = None):
    """Join the torchrun rendezvous if there is one; returns (rank, world_size)."""
    ws = env_world_size()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        # Failure detection (the reference blocks forever on a dead peer): every collective / control
        # message carries this timeout, so a rank whose peer died raises instead of hanging the job.
        import datetime
        timeout = datetime.timedelta(seconds=float(os.environ.get("FLUTE_COMM_TIMEOUT_S", "1800")))
        if backend == "nccl":
            torch.cuda.set_device(env_local_rank())
            dist.init_process_group(backend="nccl", rank=env_rank(), world_size=ws, timeout=timeout,
                                    device_id=torch.device("cuda", env_local_rank()))
        else:
            dist.init_process_group(backend=backend, rank=env_rank(), world_size=ws, timeout=timeout)
    elif torch.cuda.is_available():
        torch.cuda.set_device(env_local_rank() if env_local_rank() < torch.cuda.device_count() else 0)
    return env_rank(), ws


def make_communicator(kind: str = "auto", device=None) -> Communicator:
    """``kind``: auto | symm | collective | p2p (p2p uses the collective transport for individual payloads)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return LocalComm(device)
    backend = dist.get_backend()
    import os
    # auto: the fused P2P transport (one kernel on the server gathers, reduces, updates and re-broadcasts through NVLink
    # peer mappings) whenever symmetric memory can be set up; NCCL broadcast/reduce is the fallback / baseline
    # (``comm: collective`` or FLUTE_COMM=collective).
    env = os.environ.get("FLUTE_COMM", "")
    if kind == "auto" and env in ("symm", "collective"):
        kind = env if env == "collective" else "auto_symm"
    want_symm = kind in ("symm", "auto", "auto_symm")
    if want_symm and backend == "nccl" and torch.cuda.is_available():
        try:
            from .symm import SymmComm
            return SymmComm(device)
        except Exception as e:  # symmetric memory unavailable (no P2P, old driver …)
            if kind == "symm":
                raise
            print_rank("symmetric-memory transport unavailable ({}); falling back to collectives".format(e))
    return CollectiveComm(device)
