"""Symmetric-memory transport: broadcast / gather over NVLink peer mappings, no NCCL on the data path.

Every rank allocates the same two flat fp32 buffers from CUDA VMM symmetric memory
(``torch.distributed._symmetric_memory``: ``cuMemCreate`` + handle exchange + peer ``cuMemMap``; used only as the
allocator/bootstrap — all data movement below is this repo's own kernels):

``w_global``  the round's global weights            — written by the SERVER's update kernel into every rank's copy
``acc``       Σ weight·(w_global − w_local) of the rank — read by the SERVER's update kernel from every rank

so one FL round needs exactly one kernel on the server for "gather + reduce + DP + optimizer + broadcast"
(``csrc/server_update.cu``: P2P ``ld.global`` from peers, P2P ``st.global`` to peers) bracketed by two device-side
barriers (signal-pad flags, enqueued on the stream — the host never blocks):

    workers: train … → [barrier A] ─────────────────────────────→ [barrier B] → zero acc → next round
    server : train … → [barrier A] → fused reduce/update/bcast kernel → [barrier B]

Roofline: ingress (W−1)·P·4 B over 770 GB/s measured per direction, egress the same; for ResNet-18
(46.8 MB) on 8 GPUs ≈ 0.43 ms each way, overlapped element-tile by element-tile inside the kernel.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from ..ops import _ext
from .comm import CollectiveComm


class SymmComm(CollectiveComm):
    kind = "symm"

    def __init__(self, device=None):
        super().__init__(device)
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.group_name = dist.group.WORLD.group_name
        try:
            symm_mem.enable_symm_mem_for_group(self.group_name)
        except Exception:
            pass
        self._bufs = {}          # name -> (local tensor, handle)
        self._peer_cache = {}    # name -> [tensor of rank 0, rank 1, ...] (peer views are created once)
        self._ext = _ext.load(required=True)
        # probe: fail early (→ collective fallback) if peer mapping is not possible on this box
        probe = self._alloc("probe", 1024, torch.float32)
        probe[0].fill_(float(self.rank))
        probe[1].barrier()
        peer = probe[1].get_buffer((self.rank + 1) % self.size, (1024,), torch.float32)
        got = float(peer[0].item())
        if got != float((self.rank + 1) % self.size):
            raise RuntimeError("symmetric memory probe read {} from peer".format(got))
        probe[1].barrier()

    def _alloc(self, name, numel, dtype):
        t = self._sm.empty(numel, dtype=dtype, device=self.device)
        t.zero_()
        hdl = self._sm.rendezvous(t, group=self.group_name)
        self._bufs[name] = (t, hdl)
        return t, hdl

    # every rank must allocate in the same order: ``w_global`` first, then ``acc`` (Worker does exactly that)
    def alloc_flat(self, numel, dtype=torch.float32, name=""):
        if name in self._bufs:
            return self._bufs[name][0]
        return self._alloc(name, numel, dtype)[0]

    def _peers(self, name) -> List[torch.Tensor]:
        cached = self._peer_cache.get(name)
        if cached is not None:
            return cached
        t, hdl = self._bufs[name]
        out = []
        for r in range(self.size):
            out.append(t if r == self.rank else hdl.get_buffer(r, (t.numel(),), t.dtype))
        self._peer_cache[name] = out
        return out

    def _handle_of(self, t):
        for name, (buf, hdl) in self._bufs.items():
            if buf.data_ptr() == t.data_ptr():
                return name, hdl
        return None, None

    def device_barrier(self, t=None):
        hdl = self._bufs["acc"][1] if "acc" in self._bufs else next(iter(self._bufs.values()))[1]
        hdl.barrier()

    # ---- data plane ------------------------------------------------------------------------------------
    def broadcast_weights(self, w, src=0):
        name, hdl = self._handle_of(w)
        if name is None:
            return super().broadcast_weights(w, src)
        if self.rank == src:
            self._ext.p2p_broadcast(w, self._peers(name))       # P2P stores into every peer's copy
            _ext.count_launch(1)
        hdl.barrier()

    def reduce_accumulators(self, acc, dst=0):
        """Barrier A.  The reduction itself happens inside the server's fused update kernel, which reads
        ``peer_accumulators``; non-server ranks additionally enqueue barrier B + the zeroing of their accumulator."""
        name, hdl = self._handle_of(acc)
        if name is None:
            return super().reduce_accumulators(acc, dst)
        hdl.barrier()
        if self.rank != dst:
            hdl.barrier()
            acc.zero_()

    def peer_accumulators(self, acc):
        name, _ = self._handle_of(acc)
        return self._peers(name) if name is not None else [acc]

    def peer_weight_buffers(self, w):
        name, _ = self._handle_of(w)
        return self._peers(name) if name is not None else None

    # ---- transport v2: every rank reduces + updates + broadcasts 1/N of the arena -------------------------
    def supports_sharded(self):
        import os
        return os.environ.get("FLUTE_SHARDED_UPDATE", "1") != "0" and hasattr(self._ext, "sharded_server_update")

    def sharded_round(self, w_buf, acc, local_wsum, opt, noise_scale=0.0, seed=0):
        """Called by EVERY rank, once per round, with the same ``opt`` (the server optimizer's hyper-parameters of this
        step): [sum(weights) -> symmetric scalar] [barrier A] [sharded kernel: P2P / NVLS reduce of my slice of every
        accumulator -> optimizer -> store of my slice into every rank's weight buffer] [barrier B] [zero my accumulator].
        No NCCL call and no single-GPU funnel on the path.  Returns the rank-0 mirrors of the optimizer state
        (``(m, v)`` symmetric buffers, or ``None``) so the server can refresh its checkpointable state."""
        import os
        name_w, hdl_w = self._handle_of(w_buf)
        name_a, hdl_a = self._handle_of(acc)
        assert name_w is not None and name_a is not None, "sharded update needs symmetric weight / accumulator buffers"
        P = w_buf.numel()
        if "wsum" not in self._bufs:                         # collective allocation: every rank is in the same round
            self._alloc("wsum", 32, torch.float32)
        wsum_t = self._bufs["wsum"][0]
        wsum_t[0:1].copy_(local_wsum.reshape(1).to(wsum_t.dtype))
        kind = int(opt["code"])
        need_m = bool(opt.get("need_m", False))
        need_v = bool(opt.get("need_v", False))
        m = v = None
        m_mirror, v_mirror = [], []
        if need_m:
            if "m_srv" not in self._bufs:
                self._alloc("m_srv", P, torch.float32)
            if getattr(self, "_m_shard", None) is None:
                self._m_shard = torch.zeros(P, device=w_buf.device)
            m = self._m_shard
            m_mirror = [self._peers("m_srv")[0]]
        if need_v:
            if "v_srv" not in self._bufs:
                self._alloc("v_srv", P, torch.float32)
            if getattr(self, "_v_shard", None) is None:
                self._v_shard = torch.zeros(P, device=w_buf.device)
            v = self._v_shard
            v_mirror = [self._peers("v_srv")[0]]
        use_nvls = os.environ.get("FLUTE_NVLS", "1") != "0"
        acc_mc = int(getattr(hdl_a, "multicast_ptr", 0) or 0) if use_nvls else 0
        w_mc = int(getattr(hdl_w, "multicast_ptr", 0) or 0) if use_nvls else 0
        if not (acc_mc and w_mc):
            acc_mc = w_mc = 0
        self.last_sharded_mode = "nvls-multimem" if acc_mc else "p2p"
        hdl_a.barrier()                                      # A: every accumulator and sum(weights) is final
        b1, b2 = opt.get("betas", (0.9, 0.999))
        self._ext.sharded_server_update(
            self._peers(name_w), self._peers(name_a), self._peers("wsum"), m, v, m_mirror, v_mirror, acc_mc, w_mc,
            self.rank, kind, int(opt["step"]), float(opt["lr"]), float(b1), float(b2), float(opt.get("eps", 1e-8)),
            float(opt.get("weight_decay", 0.0)), float(opt.get("momentum", 0.0)), float(opt.get("dampening", 0.0)),
            bool(opt.get("nesterov", False)), bool(opt.get("correct_bias", True)), float(noise_scale), int(seed))
        _ext.count_launch(1)
        hdl_a.barrier()                                      # B: every slice is everywhere
        acc.zero_()
        if self.rank == 0 and (need_m or need_v):
            return (self._bufs["m_srv"][0] if need_m else None, self._bufs["v_srv"][0] if need_v else None)
        return None

    def round_done(self, acc):
        """Barrier B on the server (after its update kernel consumed every accumulator and wrote every weight copy)."""
        name, hdl = self._handle_of(acc)
        if hdl is not None:
            hdl.barrier()
