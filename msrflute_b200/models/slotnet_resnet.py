"""SlotNet executor for the GroupNorm ResNet family — the flagship client step as a static program of TMA-fed tcgen05
kernels (``csrc/slotnet.cu``).

This module is the "compiler": it walks the template ``RESNET`` (BasicBlock variants, GroupNorm with 2 channels per
group) once, lays out

* the per-slot parameter arena: every convolution filter as ``[Cout, live taps, Cin]`` (taps that only ever multiply
  zero padding are not stored — ``SlotBatchedResNet.plan_compact`` explains why that is exact), the stem filter as the
  ``[64, 160]`` matrix of its explicit im2col GEMM; an ``index_map`` ties every slot element to its position in the
  global (PyTorch-layout, checkpoint-compatible) arena so broadcast / gather are one gather / scatter kernel;
* NHWC activation, pre-norm, statistics and gradient buffers for every layer (allocated once — tensor maps hold
  absolute addresses);
* one ``GemmP`` record + TMA tensor maps per launch;

and ``step()`` replays the program: ~26 launches forward, ~50 backward for ResNet-18 (the round-1 executor needed
~200), no autograd, no allocator traffic.  Reference semantics: ``/root/reference/experiments/
cv_resnet_fedcifar100/model.py:115-191`` + ``group_normalization.py:59-84`` (per-group affine).
"""
from __future__ import annotations

import math

import torch

from ..ops import _ext
from .resnet_gn import BasicBlock, GroupNorm2d, RESNET

FPROP, DGRAD, WGRAD = 0, 1, 2
E_STORE, E_GNFWD, E_GNBWD, E_WGRAD = 0, 1, 2, 3


def _live_taps(Hi, Wi, KH, KW, stride, pad):
    Ho, Wo = (Hi + 2 * pad - KH) // stride + 1, (Wi + 2 * pad - KW) // stride + 1
    rows = [kh for kh in range(KH) if any(0 <= oh * stride - pad + kh < Hi for oh in range(Ho))]
    cols = [kw for kw in range(KW) if any(0 <= ow * stride - pad + kw < Wi for ow in range(Wo))]
    return [(kh, kw) for kh in rows for kw in cols]


def _pick_tn(n, other_tiles=10 ** 9, target=96):
    """N-tile width.  These layers are bound by the per-SM TMA row rate and per-CTA latency, not by the MMA: the weight
    (B operand) traffic does not depend on the tile width while the A operand is re-read once per N tile.  Prefer the
    widest tile that still gives ~one CTA per SM (``other_tiles`` = row tiles x slots; two CTAs on one SM share its TMA
    unit, so more than 148 CTAs buys nothing) and fall back to narrow tiles when the layer has few rows."""
    cands = [t for t in (256, 128, 64, 32) if t <= max(32, n)]
    for t in cands:
        if other_tiles * math.ceil(n / t) >= target:
            return t
    return cands[-1]


class _Conv:
    """Static description of one convolution (or the FC layer as a 1x1 convolution on a 1x1 map)."""

    def __init__(self, name, Cin, Cout, KH, KW, stride, pad, Hi, Wi):
        self.name, self.Cin, self.Cout, self.KH, self.KW, self.stride, self.pad = name, Cin, Cout, KH, KW, stride, pad
        self.Hi, self.Wi = Hi, Wi
        self.Ho, self.Wo = (Hi + 2 * pad - KH) // stride + 1, (Wi + 2 * pad - KW) // stride + 1
        self.taps = _live_taps(Hi, Wi, KH, KW, stride, pad)
        self.w_off = None            # offset of the [Cout, nt, Cin] filter inside a slot row

    @property
    def nt(self):
        return len(self.taps)


class SlotProgramBuilder:
    """Emitters for ``SlotProgram`` launches over ``[S, B, H, W, C]`` buffers and per-slot ``[S, P]`` arenas: tensor-map
    construction, tile geometry and the fprop / dgrad / wgrad records.  ``SlotNetResNet`` is the ResNet compiler on
    top of it; ``tests/test_slotnet_gpu.py`` builds single-layer programs from it to test each GEMM form in
    isolation."""

    #: debugging hook: ``[lbo, sbo, kstep, layout]`` forced on every MN-major operand descriptor (tools/debug_slotnet.py)
    MN_DESC = None
    #: measurement hooks forwarded to every GEMM launch (tools/profile_slotnet.py): accumulators per tile, debug bits
    NACC = None
    DBG = 0

    def __init__(self, W: torch.Tensor, G: torch.Tensor, batch: int, eps: float = 1e-5):
        ext = _ext.load(required=True)
        self.W, self.G = W, G
        self.S, self.P = int(W.shape[0]), int(W.shape[1])
        self.B = int(batch)
        self.dev = W.device
        self.eps = float(eps)
        self.prog = ext.SlotProgram()
        self._keep = []
        self.op_names = []           # one label per program op (tools/profile_slotnet.py)
        self.op_rw = {}              # op index -> (buffers read, buffers written): phase assignment of the fused kernel

    def _buf(self, shape):
        t = torch.zeros(shape, dtype=torch.float32, device=self.dev)
        self._keep.append(t)
        return t

    def conv(self, name, Cin, Cout, KH, KW, stride, pad, Hi, Wi, w_off):
        cv = _Conv(name, Cin, Cout, KH, KW, stride, pad, Hi, Wi)
        cv.w_off = w_off
        return cv

    def run(self, begin=0, end=-1):
        n = self.prog.run(begin, end)
        _ext.count_launch(int(n))
        return n

    # ---- tensor maps ---------------------------------------------------------------------------------------------
    def _act_map(self, t, box, parity=None, mn=False):
        S, B, H, W_, C = t.shape
        es = 4
        ptr = t.data_ptr()
        if parity is None:
            dims, strides = [C, W_, H, B, S], [C * es, W_ * C * es, H * W_ * C * es, B * H * W_ * C * es]
        else:
            py, px = parity
            ptr += (py * W_ + px) * C * es
            dims = [C, W_ // 2, H // 2, B, S]
            strides = [2 * C * es, 2 * W_ * C * es, H * W_ * C * es, B * H * W_ * C * es]
        return self.prog.add_map(ptr, dims, strides, list(box), 1 if mn else 0)

    def _w_map(self, cv, box, Cin=None, mn=False):
        """``mn``: the tile is consumed as an MN-major operand (TMA swizzle 128B_ATOM_32B, see csrc/slotnet.cu)."""
        Cin = cv.Cin if Cin is None else Cin
        ptr = self.W.data_ptr() + cv.w_off * 4
        return self.prog.add_map(ptr, [Cin, cv.nt, cv.Cout, self.S], [Cin * 4, cv.nt * Cin * 4, self.P * 4], list(box),
                                 1 if mn else 0)

    def _row_geom(self, H, W_):
        """TMA row box for an H x W row space: whole images per tile when they fit in 128 rows."""
        HW = H * W_
        if HW <= 128:
            bb = max(1, min(128 // HW, self.B))
            return {"bw": W_, "bh": H, "bb": bb, "tiles_y": 1, "row_tiles": math.ceil(self.B / bb)}
        bh = 128 // W_
        return {"bw": W_, "bh": bh, "bb": 1, "tiles_y": H // bh, "row_tiles": self.B * (H // bh)}

    def _k_geom(self, H, W_):
        """32-pixel K chunks of an H x W pixel space (wgrad)."""
        HW = H * W_
        if HW <= 32:
            kbb = 32 // HW
            return {"kbw": W_, "kbh": H, "kbb": kbb, "kH": H, "kchunks": math.ceil(self.B / kbb)}
        kbh = 32 // W_
        return {"kbw": W_, "kbh": kbh, "kbb": 1, "kH": H, "kchunks": self.B * (H // kbh)}

    def _common(self):
        d = {"S": self.S, "B": self.B, "Warena": self.W.data_ptr(), "Garena": self.G.data_ptr(),
             "arena_stride": self.P, "eps": self.eps, "dbg": int(self.DBG)}
        if self.NACC is not None:
            d["nacc"] = int(self.NACC)
        return d

    @staticmethod
    def _fprop_taps(cv):
        dx, dy, mp = [], [], []
        for kh, kw in cv.taps:
            a, b = kh - cv.pad, kw - cv.pad
            if cv.stride == 1:
                dy.append(a); dx.append(b); mp.append(0)
            else:
                dy.append(a // 2); dx.append(b // 2); mp.append((a % 2) * 2 + (b % 2))
        return dx, dy, mp

    # ---- op emitters -----------------------------------------------------------------------------------------------
    def _fprop(self, cv, x, epi, Cin=None, **kw):
        """x: input activation [S,B,Hi,Wi,Cin] → rows = output pixels."""
        Cin = cv.Cin if Cin is None else Cin
        geo = self._row_geom(cv.Ho, cv.Wo)
        TN = _pick_tn(cv.Cout, geo["row_tiles"] * self.S)
        box = (32, geo["bw"], geo["bh"], geo["bb"], 1)
        if cv.stride == 1:
            maps = [self._act_map(x, box), -1, -1, -1]
        else:
            maps = [self._act_map(x, box, parity=(p >> 1, p & 1)) for p in range(4)]
        maps.append(self._w_map(cv, (32, 1, TN, 1), Cin))
        dx, dy, mp = self._fprop_taps(cv)
        d = dict(self._common(), mode=FPROP, epi=epi, H=cv.Ho, W=cv.Wo, ntaps=cv.nt, Cred=Cin, TN=TN, N=cv.Cout,
                 tap_dx=dx, tap_dy=dy, tap_map=mp, tap_w=list(range(cv.nt)), maps=maps, **geo)
        d.update({k: (v.data_ptr() if torch.is_tensor(v) else v) for k, v in kw.items() if v is not None})
        self.prog.add_gemm(d)
        self._track([x, kw.get("res")], [kw.get("out"), kw.get("out2"), kw.get("stats") if epi == E_GNFWD else None])
        self._label("fprop {} e{} TN{} grid {}x{}x{} its {}".format(
            cv.name, epi, TN, geo["row_tiles"], math.ceil(cv.Cout / TN), self.S, cv.nt * math.ceil(Cin / 32)))

    def _label(self, text):
        while len(self.op_names) < self.prog.num_ops() - 1:
            self.op_names.append("op")
        self.op_names.append(text)

    def _track(self, reads, writes):
        """Record the activation buffers the op just added reads / writes (the parameter arena is read-only during a
        step and the gradient arena only receives commutative atomics: neither orders ops)."""
        ptr = lambda t: t.data_ptr() if torch.is_tensor(t) else int(t)
        self.op_rw[self.prog.num_ops() - 1] = ({ptr(t) for t in reads if t is not None},
                                               {ptr(t) for t in writes if t is not None})

    def phases(self, begin, end):
        """Dependency level of every op in [begin, end): one more than the latest earlier op it conflicts with
        (read-after-write, write-after-write or write-after-read on an activation buffer)."""
        ph = []
        for i in range(begin, end):
            ri, wi = self.op_rw[i]
            lvl = 0
            for j in range(begin, i):
                rj, wj = self.op_rw[j]
                if (wj & ri) or (wj & wi) or (rj & wi):
                    lvl = max(lvl, ph[j - begin] + 1)
            ph.append(lvl)
        return ph

    def _dgrad(self, cv, dy_t, epi, **kw):
        """dy_t: gradient wrt the conv output [S,B,Ho,Wo,Cout] → rows = input pixels (per parity class if stride 2)."""
        if cv.stride == 1:
            geo = self._row_geom(cv.Hi, cv.Wi)
            dx = [cv.pad - kw_ for _, kw_ in cv.taps]
            dy = [cv.pad - kh for kh, _ in cv.taps]
            extra = {"ncls": 1, "ntaps": cv.nt, "H": cv.Hi, "W": cv.Wi, "tap_w": list(range(cv.nt))}
        else:
            geo = self._row_geom(cv.Ho, cv.Wo)                      # class grid == output grid (even input sizes)
            order, dx, dy = [], [], []
            cls_tap0, cls_nt = [], []
            for c in range(4):
                py, px = c >> 1, c & 1
                cls_tap0.append(len(order))
                for ti, (kh, kw_) in enumerate(cv.taps):
                    a, b = kh - cv.pad, kw_ - cv.pad
                    if a % 2 == py and b % 2 == px:
                        order.append(ti); dy.append(-(a // 2)); dx.append(-(b // 2))
                cls_nt.append(len(order) - cls_tap0[-1])
            assert all(n > 0 for n in cls_nt), "every parity class needs a tap (3x3 / stride 2 / pad 1)"
            extra = {"ncls": 4, "cls_py": [0, 0, 1, 1], "cls_px": [0, 1, 0, 1], "cls_tap0": cls_tap0, "cls_nt": cls_nt,
                     "H": cv.Ho, "W": cv.Wo, "outH": cv.Hi, "outW": cv.Wi, "tap_w": order, "ntaps": len(order)}
        TN = _pick_tn(cv.Cin, geo["row_tiles"] * extra["ncls"] * self.S)
        box = (32, geo["bw"], geo["bh"], geo["bb"], 1)
        maps = [self._act_map(dy_t, box), -1, -1, -1, self._w_map(cv, (32, 1, 32, 1), mn=True)]
        d = dict(self._common(), mode=DGRAD, epi=epi, Cred=cv.Cout, TN=TN, N=cv.Cin, tap_dx=dx, tap_dy=dy,
                 tap_map=[0] * len(dx), maps=maps, **geo)
        d.update(extra)
        if self.MN_DESC is not None:
            d["b_desc"] = list(self.MN_DESC)
        d.update({k: (v.data_ptr() if torch.is_tensor(v) else v) for k, v in kw.items() if v is not None})
        self.prog.add_gemm(d)
        self._track([dy_t, kw.get("res"), kw.get("yprev"), kw.get("zprev"), kw.get("stats")], [kw.get("out"), kw.get("out2")])
        self._label("dgrad {} e{} TN{} grid {}x{}x{} its {}".format(
            cv.name, epi, TN, geo["row_tiles"] * extra["ncls"], math.ceil(cv.Cin / TN), self.S,
            max(extra.get("cls_nt", [extra["ntaps"]])) * math.ceil(cv.Cout / 32)))

    def _wgrad(self, cv, x, dy_t, Cin=None):
        Cin = cv.Cin if Cin is None else Cin
        kg = self._k_geom(cv.Ho, cv.Wo)
        # narrow tiles (<= 64): ~96 KB of shared memory per CTA so weight-gradient launches (side stream) co-reside
        # with the data-gradient chain, and a short epilogue (a thread stores TN values)
        TN = min(64, _pick_tn(cv.Cout, math.ceil(cv.nt * Cin / 128) * kg["kchunks"] * self.S))
        box = (32, kg["kbw"], kg["kbh"], kg["kbb"], 1)
        if cv.stride == 1:
            maps = [self._act_map(x, box, mn=True), -1, -1, -1]
        else:
            maps = [self._act_map(x, box, parity=(p >> 1, p & 1), mn=True) for p in range(4)]
        maps.append(self._act_map(dy_t, box, mn=True))
        dx, dy, mp = self._fprop_taps(cv)
        Kw = cv.nt * Cin
        row_blocks = math.ceil(Kw / 128)
        tiles = row_blocks * math.ceil(cv.Cout / TN) * self.S
        ksplit = max(1, min(kg["kchunks"], 8, round(296 / max(tiles, 1))))
        d = dict(self._common(), mode=WGRAD, epi=E_WGRAD, TN=TN, N=cv.Cout, Cin=Cin, Kw=Kw, ksplit=ksplit,
                 row_blocks=row_blocks, tap_dx=dx, tap_dy=dy, tap_map=mp, ntaps=cv.nt, maps=maps,
                 out=self.G.data_ptr() + cv.w_off * 4, side=1, **kg)
        if self.MN_DESC is not None:
            d["a_desc"], d["b_desc"] = list(self.MN_DESC), list(self.MN_DESC)
        self.prog.add_gemm(d)
        self._track([x, dy_t], [])
        self._label("wgrad {} TN{} grid {}x{}x{} its {}".format(
            cv.name, TN, row_blocks * ksplit, math.ceil(cv.Cout / TN), self.S, math.ceil(kg["kchunks"] / ksplit)))


class SlotNetResNet(SlotProgramBuilder):
    STEM_K = 160                     # 7*7*3 = 147 im2col columns, padded to 5 x 32

    @staticmethod
    def supports(model, example_input=None) -> bool:
        if not isinstance(model, RESNET) or next(model.parameters()).dtype != torch.float32:
            return False
        if getattr(model, "compute_dtype", "fp32") != "fp32":
            return False
        net = model.net
        for m in net.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                return False
            if isinstance(m, GroupNorm2d) and m.channels_per_group != 2:
                return False
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            if not all(isinstance(b, BasicBlock) for b in layer):
                return False
        if example_input is not None and tuple(example_input.shape[-3:]) != (3, 32, 32):
            return False
        ext = _ext.load()
        return ext is not None and hasattr(ext, "SlotProgram")

    # ------------------------------------------------------------------------------------------------- planning
    @staticmethod
    def _convs(template: RESNET):
        """Ordered conv descriptions + block structure for a 3x32x32 input."""
        net = template.net
        convs = {}
        stem = _Conv("conv1", 3, 64, 7, 7, 2, 3, 32, 32)
        convs["conv1"] = stem
        H = (stem.Ho + 2 - 3) // 2 + 1               # max-pool 3x3 / 2 / pad 1
        blocks = []
        C = 64
        for li, layer in enumerate((net.layer1, net.layer2, net.layer3, net.layer4), start=1):
            for bi, blk in enumerate(layer):
                pre = "layer{}.{}".format(li, bi)
                s = blk.conv1.stride[0]
                Cout = blk.conv1.out_channels
                c1 = _Conv(pre + ".conv1", C, Cout, 3, 3, s, 1, H, H)
                c2 = _Conv(pre + ".conv2", Cout, Cout, 3, 3, 1, 1, c1.Ho, c1.Wo)
                ds = None
                if blk.downsample is not None:
                    dsc = getattr(blk.downsample, "0")
                    ds = _Conv(pre + ".downsample.0", C, Cout, 1, 1, dsc.stride[0], 0, H, H)
                    convs[ds.name] = ds
                convs[c1.name], convs[c2.name] = c1, c2
                blocks.append({"prefix": pre, "c1": c1, "c2": c2, "ds": ds, "Cin": C, "Cout": Cout, "Hi": H, "Ho": c1.Ho})
                C, H = Cout, c1.Ho
        fc = _Conv("fc", C, net.fc.out_features, 1, 1, 1, 0, 1, 1)
        assert H == 1, "SlotNet expects the last stage to be 1x1 (32x32 inputs)"
        convs["fc"] = fc
        return convs, blocks

    @staticmethod
    def plan(template: RESNET, layout):
        """Slot-arena layout: ``{offsets, numel, index_map, convs, blocks}``; ``index_map[j]`` = position of slot
        element j in the global arena (-1 = padding)."""
        convs, blocks = SlotNetResNet._convs(template)
        names = [n for n, _ in template.named_parameters()]
        offsets, parts, cur = {}, [], 0
        for n, o, k, sh in zip(names, layout.offsets, layout.sizes, layout.shapes):
            mod = n[len("net."):-len(".weight")] if n.startswith("net.") and n.endswith(".weight") else None
            if mod == "conv1":
                K = SlotNetResNet.STEM_K
                idx = torch.full((64, K), -1, dtype=torch.int64)
                for kh in range(7):
                    for kw in range(7):
                        for c in range(3):
                            col = (kh * 7 + kw) * 3 + c
                            idx[:, col] = o + torch.arange(64) * 147 + c * 49 + kh * 7 + kw
                idx = idx.reshape(-1)
            elif mod in convs and len(sh) == 4:
                cv = convs[mod]
                co = torch.arange(cv.Cout, dtype=torch.int64).view(-1, 1, 1)
                tap = torch.tensor([kh * cv.KW + kw for kh, kw in cv.taps], dtype=torch.int64).view(1, -1, 1)
                ci = torch.arange(cv.Cin, dtype=torch.int64).view(1, 1, -1)
                idx = (o + (co * cv.Cin + ci) * (cv.KH * cv.KW) + tap).reshape(-1)
            else:
                idx = torch.arange(o, o + k, dtype=torch.int64)
            pad = (-cur) % 32
            if pad:
                parts.append(torch.full((pad,), -1, dtype=torch.int64))
                cur += pad
            offsets[n] = cur
            parts.append(idx)
            cur += idx.numel()
        pad = (-cur) % 32
        if pad:
            parts.append(torch.full((pad,), -1, dtype=torch.int64))
            cur += pad
        return {"offsets": offsets, "numel": cur, "index_map": torch.cat(parts).to(torch.int32), "compact": [],
                "kind": "slotnet"}

    # ------------------------------------------------------------------------------------------------- build
    def __init__(self, template: RESNET, W: torch.Tensor, G: torch.Tensor, plan: dict, batch: int):
        super().__init__(W, G, batch, eps=float(template.net.bn1.eps))
        self.off = dict(plan["offsets"])
        self.convs, self.blocks = self._convs(template)
        for name, cv in self.convs.items():
            cv.w_off = self.off["net." + name + ".weight"]
        S, B = self.S, self.B
        f = lambda *shape: self._buf(shape)
        # ---- buffers -------------------------------------------------------------------------------------------
        self.xin = f(S * B, 3, 32, 32)                       # transformed input batch (NCHW float)
        self.labels = torch.zeros(S * B, dtype=torch.int64, device=self.dev)
        self.loss = torch.zeros(S, device=self.dev)
        stem = self.convs["conv1"]
        self.col = f(S, B, 16, 16, self.STEM_K)
        self.z_stem, self.dz_stem = f(S, B, 16, 16, 64), f(S, B, 16, 16, 64)
        self.st_stem = f(S * B, 32, 2)
        self.pool = f(S, B, 8, 8, 64)
        self.arg = torch.zeros(S * B * 64 * 64, dtype=torch.uint8, device=self.dev)
        self.dpool = f(S, B, 8, 8, 64)
        for blk in self.blocks:
            Ho, Co = blk["Ho"], blk["Cout"]
            for k in ("z1", "a1", "dz1", "z2", "out", "dz2", "T"):
                blk[k] = f(S, B, Ho, Ho, Co)
            blk["st1"], blk["st2"] = f(S * B, Co // 2, 2), f(S * B, Co // 2, 2)
            if blk["ds"] is not None:
                for k in ("zds", "res", "dzds"):
                    blk[k] = f(S, B, Ho, Ho, Co)
                blk["stds"] = f(S * B, Co // 2, 2)
                blk["dxds"] = f(S, B, Ho, Ho, blk["Cin"])            # ds dgrad on the half-resolution grid
                blk["dxm"] = f(S, B, blk["Hi"], blk["Hi"], blk["Cin"])   # conv1 dgrad (full resolution)
        fc = self.convs["fc"]
        self.logits, self.dlogits = f(S, B, 1, 1, fc.Cout), f(S, B, 1, 1, fc.Cout)
        self._emit_forward()
        self.n_fwd = self.prog.num_ops()
        self._emit_backward()
        self.prog.finalize()
        import os as _os
        self.prog.set_side_stream(_os.environ.get("FLUTE_SLOTNET_SIDE", "1") == "1")
        self.prog.set_pdl(_os.environ.get("FLUTE_SLOTNET_PDL", "1") == "1")
        # persistent step kernels: the 20 forward GEMMs and the whole backward pass become one cooperative launch each
        self.fused = _os.environ.get("FLUTE_SLOTNET_FUSED", "0") == "1"
        if self.fused:
            cps = int(_os.environ.get("FLUTE_SLOTNET_FUSED_CTAS", "0"))
            for b, e in ((self.fwd_fuse_begin, self.fwd_fuse_end), (self.bwd_fuse_begin, self.bwd_fuse_end)):
                self.prog.add_mega(b, e, self.phases(b, e), cps)
        self.n_ops = self.prog.num_ops()

    def _gn(self, name):
        return {"gamma_off": self.off["net." + name + ".weight"], "beta_off": self.off["net." + name + ".bias"]}

    # ---- the program -----------------------------------------------------------------------------------------------
    def _emit_forward(self):
        p, S, B = self.prog, self.S, self.B
        stem, fc = self.convs["conv1"], self.convs["fc"]
        p.add_zero(self.loss.data_ptr(), self.loss.numel() * 4)
        p.add_im2col(self.xin.data_ptr(), list(self.xin.stride()), self.col.data_ptr(), S * B)
        # stem conv = 1-tap GEMM over the im2col matrix
        self.stem_gemm = _Conv("conv1", self.STEM_K, 64, 1, 1, 1, 0, 16, 16)
        self.stem_gemm.w_off = stem.w_off
        self._fprop(self.stem_gemm, self.col, E_STORE, out=self.z_stem)
        p.add_stem(False, dict(self._gn("bn1"), z=self.z_stem.data_ptr(), stats=self.st_stem.data_ptr(),
                               pooled=self.pool.data_ptr(), arg=self.arg.data_ptr(), Warena=self.W.data_ptr(),
                               Garena=self.G.data_ptr(), arena_stride=self.P, B=B, eps=self.eps, N=S * B))
        x = self.pool
        self.fwd_fuse_begin = p.num_ops()
        for blk in self.blocks:
            pre = blk["prefix"]
            self._fprop(blk["c1"], x, E_GNFWD, out=blk["a1"], out2=blk["z1"], stats=blk["st1"], relu=1,
                        **self._gn(pre + ".bn1"))
            res = x
            if blk["ds"] is not None:
                self._fprop(blk["ds"], x, E_GNFWD, out=blk["res"], out2=blk["zds"], stats=blk["stds"], relu=0,
                            **self._gn(pre + ".downsample.1"))
                res = blk["res"]
            self._fprop(blk["c2"], blk["a1"], E_GNFWD, out=blk["out"], out2=blk["z2"], stats=blk["st2"], relu=1, res=res,
                        **self._gn(pre + ".bn2"))
            x = blk["out"]
        self._fprop(fc, x, E_STORE, out=self.logits, bias_off=self.off["net.fc.bias"])
        self.fwd_fuse_end = p.num_ops()
        p.add_ce(dict(logits=self.logits.data_ptr(), labels=self.labels.data_ptr(), dlogits=self.dlogits.data_ptr(),
                      loss=self.loss.data_ptr(), Garena=self.G.data_ptr(), arena_stride=self.P,
                      bias_off=self.off["net.fc.bias"], B=B, C=fc.Cout, rows=S * B))

    def _emit_backward(self):
        p, S, B = self.prog, self.S, self.B
        fc = self.convs["fc"]
        blocks = self.blocks
        last = blocks[-1]
        # FC: weight gradient, then data gradient fused with ReLU mask + GroupNorm backward of the last block's bn2
        self.bwd_fuse_begin = p.num_ops()
        self._wgrad(fc, last["out"], self.dlogits)
        self._dgrad(fc, self.dlogits, E_GNBWD, out=last["dz2"], out2=last["T"], stats=last["st2"], yprev=last["out"],
                    zprev=last["z2"], relu=1, **self._gn(last["prefix"] + ".bn2"))
        for i in range(len(blocks) - 1, -1, -1):
            blk = blocks[i]
            pre = blk["prefix"]
            x_in = blocks[i - 1]["out"] if i > 0 else self.pool
            # conv2: wgrad + dgrad (→ ReLU mask of a1, GroupNorm backward of bn1 → dz1)
            self._wgrad(blk["c2"], blk["a1"], blk["dz2"])
            self._dgrad(blk["c2"], blk["dz2"], E_GNBWD, out=blk["dz1"], stats=blk["st1"], yprev=blk["a1"],
                        zprev=blk["z1"], relu=1, **self._gn(pre + ".bn1"))
            self._wgrad(blk["c1"], x_in, blk["dz1"])
            prev = blocks[i - 1] if i > 0 else None
            if blk["ds"] is None:
                if prev is not None:
                    # identity skip: d out_{i-1} = dgrad(conv1) + T_i → mask → GroupNorm backward of prev.bn2
                    self._dgrad(blk["c1"], blk["dz1"], E_GNBWD, out=prev["dz2"], out2=prev["T"], stats=prev["st2"],
                                yprev=prev["out"], zprev=prev["z2"], relu=1, res=blk["T"],
                                **self._gn(prev["prefix"] + ".bn2"))
                else:
                    self._dgrad(blk["c1"], blk["dz1"], E_STORE, out=self.dpool, res=blk["T"])
            else:
                # down-sample branch: GroupNorm backward of ds.1 on T_i, 1x1/stride-2 conv wgrad + dgrad (half grid)
                p.add_gn_bwd(dict(self._gn(pre + ".downsample.1"), din=blk["T"].data_ptr(), z=blk["zds"].data_ptr(),
                                  stats=blk["stds"].data_ptr(), dz=blk["dzds"].data_ptr(), Warena=self.W.data_ptr(),
                                  Garena=self.G.data_ptr(), arena_stride=self.P, N=S * B, B=B, H=blk["Ho"], W=blk["Ho"],
                                  C=blk["Cout"]))
                self._track([blk["T"], blk["zds"], blk["stds"]], [blk["dzds"]])
                self._wgrad(blk["ds"], x_in, blk["dzds"])
                ds_half = _Conv(blk["ds"].name, blk["Cin"], blk["Cout"], 1, 1, 1, 0, blk["Ho"], blk["Ho"])
                ds_half.w_off = blk["ds"].w_off
                self._dgrad(ds_half, blk["dzds"], E_STORE, out=blk["dxds"])
                self._dgrad(blk["c1"], blk["dz1"], E_STORE, out=blk["dxm"])
                assert prev is not None
                p.add_gn_bwd(dict(self._gn(prev["prefix"] + ".bn2"), din=blk["dxm"].data_ptr(),
                                  add2=blk["dxds"].data_ptr(), y=prev["out"].data_ptr(), z=prev["z2"].data_ptr(),
                                  stats=prev["st2"].data_ptr(), tm=prev["T"].data_ptr(), dz=prev["dz2"].data_ptr(),
                                  Warena=self.W.data_ptr(), Garena=self.G.data_ptr(), arena_stride=self.P, N=S * B, B=B,
                                  H=blk["Hi"], W=blk["Hi"], C=blk["Cin"]))
                self._track([blk["dxm"], blk["dxds"], prev["out"], prev["z2"], prev["st2"]], [prev["T"], prev["dz2"]])
        self.bwd_fuse_end = p.num_ops()
        p.add_stem(True, dict(self._gn("bn1"), z=self.z_stem.data_ptr(), stats=self.st_stem.data_ptr(),
                              pooled=self.pool.data_ptr(), arg=self.arg.data_ptr(), dpool=self.dpool.data_ptr(),
                              dz=self.dz_stem.data_ptr(), Warena=self.W.data_ptr(), Garena=self.G.data_ptr(),
                              arena_stride=self.P, B=B, eps=self.eps, N=S * B))
        self._wgrad(self.stem_gemm, self.col, self.dz_stem)

    # ------------------------------------------------------------------------------------------------- run
    def step(self, x, y):
        """x: [S*B, 3, 32, 32] float (any strides) or [S, B, ...]; y: labels.  Runs forward + backward; weight
        gradients are accumulated into the gradient arena.  Returns the per-slot mean loss tensor (a view that is
        overwritten by the next step)."""
        self.xin.copy_(x.reshape(self.xin.shape))
        self.labels.copy_(y.reshape(-1))
        n = self.prog.run(0, -1)
        _ext.count_launch(int(n))
        return self.loss

    def forward_only(self, x, y):
        self.xin.copy_(x.reshape(self.xin.shape))
        self.labels.copy_(y.reshape(-1))
        n = self.prog.run(0, self.n_fwd)
        _ext.count_launch(int(n))
        return self.logits.view(self.S, self.B, -1), self.loss
