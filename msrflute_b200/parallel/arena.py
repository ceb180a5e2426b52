"""Flat parameter arenas.

Every model replica in this framework keeps its parameters (and gradients,
optimizer moments, pseudo-gradient accumulators …) in ONE contiguous fp32
buffer; the ``nn.Parameter`` objects are views into it.  That is what lets the
hot paths of an FL round be single kernels over ``[P]`` (or ``[S, P]`` for S
concurrently simulated clients) instead of per-tensor loops:

* client step  = clip + sufficient-stats + SGD in one pass   
* pseudo-grad  = ``acc += weight·(w_global − w_local)``        (K13, K17)
* server step  = reduce over ranks + /Σw + DP noise + optimizer + broadcast
                 (K16-K18, K20-K22)

and what lets the weight broadcast / gradient gather address *one* symmetric
region per rank.  The reference instead ships every tensor separately
(``core/federated.py:112-124``: 1+3n messages per model copy).

Each tensor starts on a 128-byte boundary (32 floats) so vectorised 16-byte
accesses and TMA descriptors over individual tensors are always legal; the
padding floats are zero and every arena kernel maps 0 → 0, so they never
perturb norms or sums.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

ALIGN_ELEMS = 32  # 128 bytes of fp32


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class ArenaLayout:
    """Offsets/shapes of a list of tensors packed into one flat buffer."""

    def __init__(self, shapes: Sequence[torch.Size], names: Optional[Sequence[str]] = None):
        self.shapes = [torch.Size(s) for s in shapes]
        self.names = list(names) if names is not None else [str(i) for i in range(len(self.shapes))]
        self.sizes = [int(s.numel()) for s in self.shapes]
        self.offsets: List[int] = []
        off = 0
        for n in self.sizes:
            self.offsets.append(off)
            off = _round_up(off + n, ALIGN_ELEMS)
        self.padded_numel = max(off, ALIGN_ELEMS)
        self.numel = sum(self.sizes)

    @classmethod
    def from_module(cls, module: torch.nn.Module) -> "ArenaLayout":
        named = list(module.named_parameters())
        return cls([p.shape for _, p in named], [n for n, _ in named])

    def segments(self, device=None) -> torch.Tensor:
        """int64 ``[n_tensors, 2]`` (offset, size) table for per-tensor kernels (LAMB/LARS, quantization)."""
        return torch.tensor(list(zip(self.offsets, self.sizes)), dtype=torch.int64, device=device).reshape(-1, 2)

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + n].view(s) for o, n, s in zip(self.offsets, self.sizes, self.shapes)]

    def __eq__(self, other):
        return isinstance(other, ArenaLayout) and self.shapes == other.shapes

    def __repr__(self):
        return "ArenaLayout(tensors={}, numel={}, padded={})".format(len(self.shapes), self.numel, self.padded_numel)


class FlatArena:
    """A flat buffer plus per-tensor views."""

    def __init__(self, layout: ArenaLayout, device=None, dtype=torch.float32, buffer: Optional[torch.Tensor] = None,
                 rows: int = 1):
        self.layout = layout
        if buffer is None:
            buffer = torch.zeros(rows * layout.padded_numel, dtype=dtype, device=device)
        assert buffer.numel() == rows * layout.padded_numel and buffer.is_contiguous()
        self.rows = rows
        self.buffer = buffer
        self.flat = buffer if rows == 1 else buffer.view(rows, layout.padded_numel)

    @property
    def device(self):
        return self.buffer.device

    def row(self, r: int) -> torch.Tensor:
        return self.flat if self.rows == 1 else self.flat[r]

    def views(self, r: int = 0) -> List[torch.Tensor]:
        return self.layout.views(self.row(r))

    def zero_(self):
        self.buffer.zero_()
        return self

    def copy_from_tensors(self, tensors: Iterable[torch.Tensor], r: int = 0):
        for v, t in zip(self.views(r), tensors):
            v.copy_(t)
        return self

    def like(self, rows: Optional[int] = None) -> "FlatArena":
        return FlatArena(self.layout, device=self.device, dtype=self.buffer.dtype, rows=rows or self.rows)


def adopt_module(module: torch.nn.Module, with_grad: bool = True, param_buffer: Optional[torch.Tensor] = None,
                 grad_buffer: Optional[torch.Tensor] = None):
    """Re-home ``module``'s parameters into a flat arena (in place).

    Returns ``(param_arena, grad_arena_or_None)``.  After this call
    ``p.data`` (and ``p.grad`` when ``with_grad``) of every parameter are views
    of the arenas, in ``module.parameters()`` order — the order the reference
    uses on the wire (``core/server.py:280``).
    """
    params = [p for _, p in module.named_parameters()]
    if not params:
        raise ValueError("module has no parameters")
    dev, dt = params[0].device, params[0].dtype
    layout = ArenaLayout.from_module(module)
    w = FlatArena(layout, device=dev, dtype=dt, buffer=param_buffer)
    for p, v in zip(params, w.views()):
        v.copy_(p.data)
        p.data = v
    g = None
    if with_grad:
        g = FlatArena(layout, device=dev, dtype=dt, buffer=grad_buffer)
        for p, v in zip(params, g.views()):
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
    module._flute_arena = (w, g)
    return w, g


def module_arena(module: torch.nn.Module):
    """Return the ``(param, grad)`` arenas of an adopted module, verifying the views are still intact.  The full
    check (every parameter still points into the arena) runs on the first call and every 64th; the calls in between
    only look at the first and last parameter — this sits on the per-round host path."""
    ar = getattr(module, "_flute_arena", None)
    if ar is None:
        return None
    w, g = ar
    base = w.buffer.data_ptr()
    esz = w.buffer.element_size()
    n = getattr(module, "_flute_arena_checks", 0)
    module._flute_arena_checks = n + 1
    ends = getattr(module, "_flute_arena_ends", None)
    if n % 64 != 0 and ends is not None:
        (p0, o0), (p1, o1) = ends
        if p0.data.data_ptr() == base + o0 * esz and p1.data.data_ptr() == base + o1 * esz:
            return ar
    params = list(module.parameters())
    for p, o in zip(params, w.layout.offsets):
        if p.data.data_ptr() != base + o * esz:
            return None  # somebody rebound p.data (e.g. load_state_dict(assign=True)); arena no longer valid
    if params:
        module._flute_arena_ends = ((params[0], w.layout.offsets[0]), (params[-1], w.layout.offsets[len(params) - 1]))
    return ar


def rebind_grads(module: torch.nn.Module):
    """Point ``p.grad`` back at the grad arena (``optimizer.zero_grad(set_to_none=True)`` unbinds them)."""
    ar = getattr(module, "_flute_arena", None)
    if ar is None or ar[1] is None:
        return
    for p, v in zip(module.parameters(), ar[1].views()):
        if p.grad is None or p.grad.data_ptr() != v.data_ptr():
            p.grad = v
