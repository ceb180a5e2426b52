"""Server ↔ worker orchestration and transport (ref. ``core/federated.py``).

Reference protocol: rank 0 serially sends COMMAND_UPDATE + lr + the model as 1+3n messages to
every worker, then hands sampled client ids to free workers one by one and polls ``irecv`` acks; each result
comes back as 1+3n more messages.  Rank 0 does not train when world_size > 1.

This implementation keeps the roles, the command set and the ``Server`` / ``Worker`` API but re-plumbs it:

* **control plane**: one small python dict per phase (``{'cmd','lr','round','assign',…}``) broadcast on a CPU
  gloo group — never a GPU sync, no 8-character metric-name limit (ref ``:37``).
* **data plane**: the model is a flat arena.  Round start = ``comm.broadcast_weights`` (skipped when the
  previous fused update already multicast the weights); round end = ``comm.reduce_accumulators`` of the
  per-rank weighted pseudo-gradient sums ("fused" mode) or one flat message per client ("individual" mode,
  needed for stale gradients / cosine dumps / RL — the reference's non-``fast_aggregation`` path).
* **placement**: every GPU trains (rank 0 included, ``server_config.b200.server_is_worker``); clients are
  assigned by longest-processing-time-first on ``num_samples`` so ranks finish together — all ranks compute
  the same assignment from the broadcast, no per-client dispatch messages or ack polling.
* world_size == 1 needs no thread / module-global hand-off (ref ``:382-402``): the server calls its worker.
"""
from __future__ import annotations

import cProfile
import logging
import os
import time
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..parallel.arena import adopt_module, module_arena
from ..parallel.comm import Communicator, LocalComm
from ..utils import print_profiler, print_rank, to_device
from .client import Client

COMMAND_UPDATE = 0
COMMAND_TRAIN = 1
COMMAND_SYNC_NODES = 9
COMMAND_TERMINATE = 10
COMMAND_TESTVAL = 11


def rank():
    return int(os.environ.get("RANK", 0))


def local_rank():
    return int(os.environ.get("LOCAL_RANK", 0))


def size():
    return int(os.environ.get("WORLD_SIZE", 1))


def encode_string(word, string_to_int=True):
    """ASCII <-> int list (the reference's fixed 8-char wire encoding of metric names, ``:27-43``).
    Kept for API parity; this transport sends names as-is, so longer names are not truncated."""
    if string_to_int:
        return list(word.ljust(8, " ").encode())
    return bytes([c for c in word if c != 32]).decode()


# ----------------------------------------------------------- p2p primitives
def _dev(t):
    return to_device(t) if dist.is_initialized() and dist.get_backend() == "nccl" else t


def _send(x, dst=0):
    """Send a python scalar / list as a tensor (ref ``:89-96``; no ``empty_cache`` per message here)."""
    dist.send(_dev(torch.as_tensor(x)), dst)


def _recv(x, src=0):
    t = _dev(torch.as_tensor(x))
    dist.recv(tensor=t, src=src)
    return t.item() if t.dim() == 0 else t.tolist()


#: bytes of gradient payload this process put on / took off the wire (individual-payload path; tests and logs)
WIRE_BYTES = {"sent": 0, "received": 0, "packed_msgs": 0}


def _send_gradients(gradients, dst, quant=None):
    """Ragged tensor list as ONE header + ONE flat payload (the reference needs 1+3n messages, ``:112-124``).

    ``quant`` (``{"bits", "lo_hi"}``, set by the quantization extension): the values are already snapped to ``2**bits``
    levels per tensor, so the payload travels PACKED — level codes (1 or 2 bytes), a 1-bit keep mask and a tiny
    (lo, width) table instead of fp32 values (``ops.quant_ops.wire_encode``; 3.6x fewer bytes at 8 bits).  The reference
    only simulates quantization and ships fp32 (``extensions/quantization/quant.py:9-50`` + ``federated.py:126-157``)."""
    import os
    shapes = [list(g.shape) for g in gradients]
    packed = (quant is not None and gradients and os.environ.get("FLUTE_PACKED_WIRE", "1") == "1"
              and int(quant["lo_hi"].shape[0]) == len(gradients))
    fmt = int(quant["bits"]) if packed else 0
    header = [fmt, len(shapes)] + [len(s) for s in shapes] + [d for s in shapes for d in s]
    _send(len(header), dst)
    _send(header, dst)
    flat = torch.cat([g.reshape(-1).float() for g in gradients]) if gradients else torch.zeros(0)
    if not packed:
        dist.send(_dev(flat.contiguous()), dst)
        WIRE_BYTES["sent"] += flat.numel() * 4
        return
    from ..ops import quant_ops
    codes, bitmap, table = quant_ops.wire_encode(flat, [g.numel() for g in gradients], quant["lo_hi"].to(flat.device), fmt)
    dist.send(_dev(table.contiguous()), dst)
    dist.send(_dev(codes.contiguous()), dst)
    dist.send(_dev(bitmap.contiguous()), dst)
    WIRE_BYTES["sent"] += table.numel() * 4 + codes.numel() * codes.element_size() + bitmap.numel()
    WIRE_BYTES["packed_msgs"] += 1


def _recv_gradients(src):
    n_hdr = _recv(0, src)
    header = _recv([0] * n_hdr, src)
    fmt, n = header[0], header[1]
    ndims = header[2:2 + n]
    dims, pos = [], 2 + n
    for nd in ndims:
        dims.append(header[pos:pos + nd])
        pos += nd
    sizes = [int(np.prod(d)) if len(d) else 1 for d in dims]
    total = sum(sizes)
    if fmt == 0:
        flat = _dev(torch.zeros(total))
        dist.recv(flat, src)
        WIRE_BYTES["received"] += total * 4
    else:
        from ..ops import quant_ops
        table = _dev(torch.zeros(n, 2))
        codes = _dev(torch.zeros(total if fmt <= 8 else 2 * total, dtype=torch.uint8))
        bitmap = _dev(torch.zeros((total + 7) // 8, dtype=torch.uint8))
        dist.recv(table, src)
        dist.recv(codes, src)
        dist.recv(bitmap, src)
        flat = quant_ops.wire_decode(codes, bitmap, table, sizes)
        WIRE_BYTES["received"] += table.numel() * 4 + codes.numel() * codes.element_size() + bitmap.numel()
        WIRE_BYTES["packed_msgs"] += 1
    out, off = [], 0
    for d, k in zip(dims, sizes):
        out.append(flat[off:off + k].view(d))
        off += k
    return out


# ------------------------------------------------------------------ runtime
class _Runtime:
    """Per-process singletons: the communicator and (when this rank trains) its worker."""
    comm: Communicator = None
    worker: "Worker" = None
    options: Dict = {}


def init_runtime(comm: Communicator, worker: Optional["Worker"] = None, **options):
    _Runtime.comm, _Runtime.worker, _Runtime.options = comm, worker, options


def get_comm() -> Communicator:
    if _Runtime.comm is None:
        _Runtime.comm = LocalComm()
    return _Runtime.comm


class WorkQueue:
    """Iterator over the clients THIS rank wins from a job-wide queue (``dispatch: work_queue``).

    Every training rank walks the same cost-sorted list; ``comm.fetch_add`` hands out the next position atomically, so
    a rank that finishes early simply takes more clients — load balancing by completion, like the reference's
    acknowledge-and-send-next loop, without a server thread in the middle.  ``taken`` records what this rank trained."""

    def __init__(self, comm, key: str, items: List[int]):
        self.comm, self.key, self.items, self.taken = comm, key, list(items), []

    def __iter__(self):
        while True:
            k = self.comm.fetch_add(self.key, 1)
            if k >= len(self.items):
                return
            self.taken.append(self.items[k])
            yield self.items[k]

    def __len__(self):
        return len(self.items)


def assign_clients(items: List[int], costs: List[float], workers: List[int], policy: str = "static_lpt",
                   speeds: Optional[dict] = None):
    """Split ``items`` over ``workers``.

    ``static_lpt``: sort by cost descending, always give the next item to the least-loaded worker (≤ 4/3 of the optimal
    makespan); ``round_robin``: i-th item → i mod W; ``dynamic``: LPT on *estimated finish time* ``load / speed`` where
    ``speeds[w]`` (cost units per second) is an exponential moving average of what each worker actually delivered in
    earlier rounds (:func:`update_worker_speeds`).  The reference balances load by handing the next client to whichever
    worker acknowledges first (``federated.py:282-410``); with one command per worker per round the same adaptation to
    slow / shared / heterogeneous GPUs comes from the measured speeds instead of per-client round trips."""
    out = {w: [] for w in workers}
    if not workers:
        return out
    if policy == "round_robin":
        for i, it in enumerate(items):
            out[workers[i % len(workers)]].append(it)
        return out
    speed = {w: 1.0 for w in workers}
    if policy == "dynamic" and speeds:
        known = [v for w, v in speeds.items() if w in speed and v > 0]
        default = sum(known) / len(known) if known else 1.0
        speed = {w: (speeds.get(w) if speeds.get(w, 0) > 0 else default) for w in workers}
    load = {w: 0.0 for w in workers}
    for it, c in sorted(zip(items, costs), key=lambda x: -x[1]):
        c = max(float(c), 1e-9)
        w = min(workers, key=lambda k: ((load[k] + c) / speed[k], k))
        out[w].append(it)
        load[w] += c
    return out


def update_worker_speeds(assign, costs_by_item, seconds_by_worker, momentum=0.7):
    """EMA of cost units per second per worker from one round's measurements (``dynamic`` dispatch)."""
    speeds = _Runtime.options.setdefault("_worker_speeds", {})
    for w, items in assign.items():
        t = seconds_by_worker.get(w, 0.0)
        if not items or t <= 0:
            continue
        v = sum(max(float(costs_by_item.get(i, 1.0)), 1e-9) for i in items) / t
        speeds[w] = v if w not in speeds else momentum * speeds[w] + (1 - momentum) * v
    return speeds


class Server:
    """Orchestration half of the server (aggregation lives in ``core/server.py``)."""

    @staticmethod
    def _workers(comm):
        opts = _Runtime.options
        if comm.size == 1:
            return [0]
        return list(range(0 if opts.get("server_is_worker", True) else 1, comm.size))

    @staticmethod
    def dispatch_clients(clients, server_data, command, mode=None, do_profiling=False, single_worker=None,
                         costs=None, fused=False, sync_weights=True, extra=None, defer=False):
        """Run ``command`` for ``clients`` on the available workers; generator of per-client outputs.

        ``server_data`` = ``(lr, weights, round)`` where ``weights`` is the server's flat arena tensor (fast path),
        a list of tensors (reference format) or None (keep what workers already hold).  ``defer`` (single-process
        fused training on the device engine) yields ONE :class:`~.engine.DeferredRound` instead of per-client
        dicts: nothing has been read back from the GPU yet.  In ``fused`` TRAIN mode the outputs carry ``pl = {'weight', 'gradients': None, 'fused': True}`` and the weighted pseudo-gradient sum
        ends up in the server worker's accumulator (see :meth:`take_accumulator`)."""
        comm = get_comm()
        worker = single_worker or _Runtime.worker
        profiler = None
        if do_profiling:
            profiler = cProfile.Profile()
            profiler.enable()
        lr, weights, nround = server_data
        workers = Server._workers(comm)
        costs = costs if costs is not None else [1.0] * len(clients)
        policy = _Runtime.options.get("dispatch", "static_lpt")
        queue = None
        if policy == "work_queue":
            # pull-based dispatch for the per-client (generic) training path; the slot engine consumes whole waves and
            # evaluation chunks are uniform, so those keep the static plan
            engine_takes_it = (worker is not None and getattr(worker, "engine", None) is not None and fused
                               and worker.engine.supports(worker.config))
            if command == COMMAND_TRAIN and comm.size > 1 and not engine_takes_it and not (defer and fused):
                order = [c for _, c in sorted(zip(costs, clients), key=lambda t: -t[0])]
                _Runtime.options["_wq_seq"] = _Runtime.options.get("_wq_seq", 0) + 1
                queue = {"key": "wq/{}/{}".format(nround, _Runtime.options["_wq_seq"]), "items": order}
            policy = "static_lpt"
        assign = assign_clients(list(clients), list(costs), workers, policy,
                                speeds=_Runtime.options.get("_worker_speeds"))
        if queue is not None:
            assign = {w: [] for w in workers}
        cost_of = dict(zip(clients, costs))
        flat_ok = torch.is_tensor(weights)
        ctrl = {"cmd": command, "lr": lr, "round": nround, "assign": assign, "mode": mode, "fused": fused,
                "sync": ("flat" if flat_ok else "list" if weights is not None else "none") if sync_weights else "none",
                "extra": extra or {}, "defer": bool(defer and fused and command == COMMAND_TRAIN), "queue": queue}
        if comm.size > 1:
            comm.bcast_object(ctrl, src=0)
            Server._sync_weights(comm, worker, weights, ctrl["sync"])
        elif worker is not None and weights is not None and ctrl["sync"] != "none":
            worker.set_weights(weights)

        # local share
        local_out = []
        my_rank = comm.rank if comm.size > 1 else 0
        if worker is not None and (assign.get(my_rank) or (queue is not None and my_rank in workers)):
            mine = WorkQueue(comm, queue["key"], queue["items"]) if queue is not None else assign[my_rank]
            if command == COMMAND_TRAIN:
                local_out = worker.train_clients(mine, (lr, None, nround), fused=fused, extra=ctrl["extra"],
                                                 defer=ctrl["defer"])
                if comm.size == 1 and not isinstance(local_out, list):      # DeferredRound
                    yield local_out
                    return
            else:
                local_out = worker.eval_clients(mine, mode, (lr, None, nround))
        if ctrl["defer"] and comm.size > 1:
            # multi-rank deferred round: the cross-rank reduction of the accumulators and of Σ weight is enqueued
            # right behind the local training; per-client records are gathered (host, gloo) only on resolve()
            yield _finish_deferred(comm, worker, local_out, len(clients), assign=assign,
                                   cost_of=cost_of if policy == "dynamic" else None,
                                   sharded=(ctrl["extra"] or {}).get("sharded"))
            return
        for o in local_out:
            yield o

        # remote shares
        if comm.size > 1:
            if command == COMMAND_TRAIN:
                records = comm.gather_objects([_strip(o) for o in local_out])
                if fused:
                    comm.reduce_accumulators(worker.accumulator(), dst=0)
                if policy == "dynamic":
                    update_worker_speeds(assign, cost_of, {r: sum(float(o["cs"].get("training", 0.0)) for o in recs)
                                                          for r, recs in enumerate(records)})
                for r, recs in enumerate(records):
                    if r == comm.rank:
                        continue
                    for rec in recs:
                        if not fused:
                            rec["pl"]["gradients"] = _recv_gradients(r)
                        yield rec
            else:
                results = comm.gather_objects(local_out)
                for r, res in enumerate(results):
                    if r != comm.rank:
                        for o in res:
                            yield o
        if do_profiling:
            profiler.disable()
            print_profiler(profiler)

    @staticmethod
    def _sync_weights(comm, worker, weights, how):
        if how == "none":
            return
        if how == "flat":
            from ..utils.timing import PHASES
            buf = worker.weight_buffer() if worker is not None else weights
            with PHASES.phase("bcast_xgpu"):
                if buf.data_ptr() != weights.data_ptr():
                    buf.copy_(weights)
                comm.broadcast_weights(buf, src=0)
            if worker is not None:
                worker.weights_updated()
        else:
            tensors = comm.bcast_object([t.cpu() for t in weights], src=0)
            if worker is not None:
                worker.set_weights(tensors)

    @staticmethod
    def process_clients(clients, server_data, single_worker=None, **kw):
        return Server.dispatch_clients(clients, server_data, COMMAND_TRAIN, single_worker=single_worker, **kw)

    @staticmethod
    def process_testvalidate(clients, server_data, mode, single_worker=None, **kw):
        return Server.dispatch_clients(clients, server_data, COMMAND_TESTVAL, mode, single_worker=single_worker, **kw)

    @staticmethod
    def sync_nodes(options=None):
        """Barrier across all ranks driven from the server (workers sit in their command loop).  Every rank drains its
        GPU first, so "barrier + synchronize" brackets of a benchmark really bracket device work.  ``options`` is
        applied on every worker before the barrier (e.g. ``{"resident": False}`` switches the engines to streaming)."""
        comm = get_comm()
        if comm.size > 1 and comm.rank == 0:
            comm.bcast_object({"cmd": COMMAND_SYNC_NODES, "sync": "none", "assign": {}, "options": options or {}}, src=0)
        worker = _Runtime.worker
        if worker is not None and options:
            worker.apply_options(options)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if comm.size > 1:
            comm.barrier()

    @staticmethod
    def terminate_workers(terminate=True):
        comm = get_comm()
        if terminate and comm.size > 1 and comm.rank == 0:
            comm.bcast_object({"cmd": COMMAND_TERMINATE}, src=0)
            if WIRE_BYTES["packed_msgs"]:
                print_rank("gradient payloads on the wire: {} packed messages, {:.2f} MB received".format(
                    WIRE_BYTES["packed_msgs"], WIRE_BYTES["received"] / 1e6), logging.INFO)


def _finish_deferred(comm, worker, local, n_clients, assign=None, cost_of=None, sharded=None):
    """Every rank calls this with its (possibly deferred / empty) local result, in the same order: NCCL/symm reduce of
    the accumulators → all-reduce of Σ weight → (on resolve) host sync + gloo gather of the per-client records."""
    from .engine import DeferredRound
    acc = worker.accumulator()
    if isinstance(local, DeferredRound):
        wsum = local.weight_sum.reshape(1).float().clone()
    else:
        wsum = torch.tensor([float(sum((o.get("pl") or {}).get("weight", 0.0) for o in local))], device=acc.device)
    # Σ weight first: with the symmetric-memory transport the server's stream is [barrier A, update kernel, barrier B]
    # and the update kernel needs the total, so the (NCCL) all-reduce must precede barrier A on every rank.
    from ..utils.timing import PHASES
    mirrors = None
    if sharded is not None:
        # transport v2: no all-reduce, no funnel — every rank reduces / updates / broadcasts its slice of the arena
        with PHASES.phase("update_bcast"):
            mirrors = comm.sharded_round(worker.weight_buffer(), acc, wsum, sharded["opt"],
                                         noise_scale=sharded.get("noise_scale", 0.0), seed=sharded.get("seed", 0))
    else:
        with PHASES.phase("gather_xgpu"):
            comm.all_reduce_(wsum)
            comm.reduce_accumulators(acc, dst=0)
            if comm.rank != 0 and comm.kind != "symm":
                acc.zero_()

    def resolve():
        outs = local.resolve() if isinstance(local, DeferredRound) else list(local)
        records = comm.gather_objects([_strip(o) for o in outs])
        merged = list(outs)
        for r, recs in enumerate(records):
            if r != comm.rank:
                merged.extend(recs)
        if cost_of is not None:                                     # ``dynamic`` dispatch: learn the workers' speeds
            update_worker_speeds(assign, cost_of, {r: sum(float(o["cs"].get("training", 0.0)) for o in recs)
                                                  for r, recs in enumerate(records)})
        return merged

    dr = DeferredRound(n_clients, wsum[0], resolve)
    dr.sharded_done = sharded is not None
    dr.state_mirrors = mirrors
    return dr


def _maybe_inject_fault(rank, nround):
    """Fault injection for failure-detection tests: ``FLUTE_FAULT_INJECT=<rank>:<round>`` makes that worker die
    (``os._exit(17)``) when it receives the command of that round."""
    spec = os.environ.get("FLUTE_FAULT_INJECT", "")
    if spec:
        r, n = spec.split(":")
        if int(r) == rank and int(n) == int(nround):
            print_rank("fault injection: rank {} exits at round {}".format(rank, nround), logging.WARNING)
            os._exit(17)


def _strip(o):
    """Client record without the (device) gradient payload — what travels on the control plane."""
    rec = {k: v for k, v in o.items() if k != "pl"}
    pl = o.get("pl") or {}
    rec["pl"] = {"weight": pl.get("weight", 0.0), "gradients": None, "fused": pl.get("fused", False)}
    for k in ("mg", "vg", "ng", "rg"):
        if k in rec:
            rec[k] = np.float32(rec[k])
    return rec


class Worker:
    """Processes simulated clients / evaluation chunks on one GPU (ref. ``federated.py:452-676``)."""

    def __init__(self, model=None, data_path=None, do_profiling=False, val_clients=None, test_clients=None,
                 config=None, val_dataset=None, test_dataset=None):
        self.model = model
        self.data_path = data_path
        self.do_profiling = do_profiling
        self.config = config
        self.val_clients, self.test_clients = val_clients, test_clients
        self.val_dataset, self.test_dataset = val_dataset, test_dataset
        self._w = None           # flat copy of the global weights for this round
        self._acc = None         # Σ weight·pseudo-gradient over this rank's clients (fused mode)
        self._weights_list = None
        self.engine = None       # optional device-resident multi-client engine (core/engine.py)
        self.round_events = []   # CUDA events recorded after every TRAIN command (device-side round timing)
        self.sync_events = []    # CUDA events recorded at every server-driven sync_nodes()

    # ---- weight / accumulator buffers -------------------------------------
    def _arena(self):
        ar = module_arena(self.model)
        if ar is None:
            self.model = to_device(self.model)
            ar = adopt_module(self.model, with_grad=True)
        return ar

    def weight_buffer(self) -> torch.Tensor:
        if self._w is None:
            w = self._arena()[0]
            self._w = get_comm().alloc_flat(w.flat.numel(), w.flat.dtype, name="w_global")
            if self._w.device != w.flat.device:
                self._w = torch.zeros_like(w.flat)
        return self._w

    def accumulator(self) -> torch.Tensor:
        if self._acc is None:
            w = self._arena()[0]
            self._acc = get_comm().alloc_flat(w.flat.numel(), w.flat.dtype, name="acc")
            if self._acc.device != w.flat.device:
                self._acc = torch.zeros_like(w.flat)
        return self._acc

    def set_weights(self, weights):
        if torch.is_tensor(weights):
            buf = self.weight_buffer()
            if buf.data_ptr() != weights.data_ptr():
                buf.copy_(weights)
            self._weights_list = None
        else:
            self._weights_list = [t for t in weights]
            lay = self._arena()[0].layout
            if len(weights) == len(lay.shapes):
                buf = self.weight_buffer()
                for v, t in zip(lay.views(buf), weights):
                    v.copy_(t)
                self._weights_list = None

    def weights_updated(self):
        self._weights_list = None

    def _server_weights(self):
        return self._weights_list if self._weights_list is not None else self.weight_buffer()

    # ---- work ---------------------------------------------------------------
    def train_clients(self, client_ids, server_data, fused=False, extra=None, defer=False):
        lr, _, nround = server_data
        cfg = self.config
        if extra:
            if extra.get("quant_thresh") is not None:
                cfg["client_config"]["quant_thresh"] = extra["quant_thresh"]
            if extra.get("max_allowed_leakage") is not None:
                cfg["privacy_metrics_config"]["max_allowed_leakage"] = extra["max_allowed_leakage"]
        if self.engine is not None and fused and self.engine.supports(cfg):
            return self.engine.train_clients(client_ids, lr, nround, self._server_weights(), self.accumulator(),
                                             defer=defer)
        outs = []
        send_grads = cfg["client_config"].get("type", "gradient_computation") == "optimization"
        profiler = None
        if self.do_profiling:
            profiler = cProfile.Profile()
            profiler.enable()
        for cid in client_ids:
            client = Client([cid], cfg, send_grads)
            out = Client.process_round(client.get_client_data(), (lr, self._server_weights(), nround), self.model,
                                       self.data_path)
            if fused:
                pl = out.get("pl") or {"weight": 0.0}
                w = 0.0 if out.get("wt", 1.0) == 0.0 else pl["weight"]
                if w != 0.0 and pl.get("flat") is not None:
                    self.accumulator().add_(pl["flat"])
                out["pl"] = {"weight": w, "gradients": None, "fused": True}
            outs.append(out)
        if self.do_profiling:
            profiler.disable()
            print_profiler(profiler)
        return outs

    def eval_clients(self, chunk_ids, mode, server_data):
        mode = mode if isinstance(mode, str) else ("test" if (mode[0] if isinstance(mode, (list, tuple)) else mode) == -2 else "val")
        clients = self.val_clients if mode == "val" else self.test_clients
        dataset = self.val_dataset if mode == "val" else self.test_dataset
        outs = []
        for idx in chunk_ids:
            c = clients[idx]
            weights = self._server_weights()
            out, metrics, n = Client.run_testvalidate(c.get_client_data(dataset), (0.0, weights, 0), mode, self.model)
            outs.append((out, metrics, n))
        return outs

    # ---- command loop for ranks that are not the server -----------------------------
    def run(self):
        comm = get_comm()
        if comm.size == 1 or comm.rank == 0:
            return
        while True:
            ctrl = comm.bcast_object(None, src=0)
            cmd = ctrl["cmd"]
            if cmd == COMMAND_TERMINATE:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                return
            if cmd == COMMAND_UPDATE:
                Server._sync_weights(comm, self, self.weight_buffer(), ctrl.get("sync", "flat"))
                continue
            if cmd == COMMAND_SYNC_NODES:
                self.apply_options(ctrl.get("options") or {})
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    self.sync_events.append(ev)
                comm.barrier()
                continue
            if ctrl["sync"] == "flat":
                comm.broadcast_weights(self.weight_buffer(), src=0)
                self.weights_updated()
            elif ctrl["sync"] == "list":
                self.set_weights(comm.bcast_object(None, src=0))
            mine = ctrl["assign"].get(comm.rank, [])
            if ctrl.get("queue") is not None and comm.rank in ctrl["assign"]:
                mine = WorkQueue(comm, ctrl["queue"]["key"], ctrl["queue"]["items"])
            _maybe_inject_fault(comm.rank, ctrl.get("round", -1))
            if cmd == COMMAND_TRAIN:
                if ctrl.get("defer"):
                    res = self.train_clients(mine, (ctrl["lr"], None, ctrl["round"]), fused=True,
                                             extra=ctrl.get("extra"), defer=True) if mine else []
                    _finish_deferred(comm, self, res, len(mine), sharded=(ctrl.get("extra") or {}).get("sharded")).resolve()
                    if torch.cuda.is_available():
                        ev = torch.cuda.Event(enable_timing=True)
                        ev.record()
                        self.round_events.append(ev)
                        del self.round_events[:-64]
                    continue
                outs = self.train_clients(mine, (ctrl["lr"], None, ctrl["round"]), fused=ctrl["fused"],
                                          extra=ctrl.get("extra"))
                if isinstance(mine, WorkQueue):
                    print_rank("work queue: rank {} trained {} of {} clients of round {}".format(
                        comm.rank, len(mine.taken), len(mine), ctrl["round"]), logging.INFO)
                comm.gather_objects([_strip(o) for o in outs])
                if ctrl["fused"]:
                    comm.reduce_accumulators(self.accumulator(), dst=0)
                    self.accumulator().zero_()
                else:
                    for o in outs:
                        pl = o.get("pl") or {}
                        _send_gradients(pl.get("gradients") or [], 0, quant=pl.get("quant"))
                if torch.cuda.is_available():
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    self.round_events.append(ev)
                    del self.round_events[:-64]
            elif cmd == COMMAND_TESTVAL:
                comm.gather_objects(self.eval_clients(mine, ctrl["mode"], (0.0, None, 0)))
            else:
                raise AssertionError("unknown command {}".format(cmd))

    def apply_options(self, options):
        if "resident" in options and self.engine is not None:
            self.engine.set_resident(bool(options["resident"]))
        if "phases" in options:                       # bench.py: per-phase device timers (exposed comm per round)
            from ..utils.timing import PHASES
            PHASES.enable(bool(options["phases"]))

    def sync_region_ms(self, first, second):
        """Device time between this rank's ``first``-th and ``second``-th ``sync_nodes`` (CUDA events)."""
        ev = self.sync_events
        if len(ev) <= max(first, second):
            return 0.0
        ev[second].synchronize()
        return float(ev[first].elapsed_time(ev[second]))

    # ---- reference single-process entry points ------------------------------------------
    def trigger_train(self, lr, model_params, nround, client_idx):
        self.set_weights(model_params)
        return self.train_clients([client_idx], (lr, None, nround))[0]

    def trigger_evaluate(self, model_params, mode="val", idx=0):
        self.set_weights(model_params)
        return self.eval_clients([idx], mode, (0.0, None, 0))[0]
