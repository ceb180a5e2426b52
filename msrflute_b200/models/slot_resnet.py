"""Slot-batched executor for the GroupNorm ResNet family (the flagship FedCIFAR-100 model).

Runs the forward/backward of S simulated clients' ResNets as ONE sequence of launches: every convolution is
``ops.slot_ops.SlotConv2d`` (hand-written direct conv over all slots, weights addressed inside the ``[S, P]`` arena),
every norm is ``SlotGroupNorm`` (norm + affine + residual + ReLU fused), the classifier is a batched GEMM.  Weight
gradients are accumulated by the backward kernels straight into the gradient arena.  The executor mirrors
``models/resnet_gn.ResNet.forward`` layer by layer using the template module only for its structure (names, strides,
group counts) — the numbers come from the arenas.
"""
import torch
import torch.nn.functional as F

from ..ops import misc_ops
from ..ops.slot_ops import SlotConv2d, SlotGroupNorm, SlotLinear, live_taps
from .resnet_gn import BasicBlock, Bottleneck, GroupNorm2d, RESNET


import os as _os

_SLOT_MAXPOOL = _os.environ.get("FLUTE_SLOT_MAXPOOL", "1") == "1"


class SlotBatchedResNet:
    @staticmethod
    def supports(model) -> bool:
        if not isinstance(model, RESNET) or next(model.parameters()).dtype != torch.float32:
            return False
        if getattr(model, "compute_dtype", "fp32") != "fp32":
            return False
        return not any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in model.modules())

    def __init__(self, template: RESNET, layout, W, G, plan=None):
        self.net = template.net
        self.W, self.G, self.S = W, G, W.shape[0]
        names = [n for n, _ in template.named_parameters()]
        self.off = dict(plan["offsets"]) if plan is not None else {n: o for n, o in zip(names, layout.offsets)}
        self.compact = set(plan["compact"]) if plan is not None else set()
        self.dummy = torch.zeros(1, device=W.device, requires_grad=True)

    @staticmethod
    def plan_compact(template: RESNET, layout, example_input):
        """Slot-arena layout that stores only the LIVE filter taps of every convolution.

        On small feature maps most taps of a 3x3/pad-1 filter only ever multiply zero padding (ResNet-18 on 32x32
        inputs: layer4 runs on 1x1 maps, 8 of 9 taps are dead = 63 % of all parameters).  Their gradient is identically
        zero, so with ``weight_decay == 0`` they never change on a client and need not exist per client: the slot arenas
        shrink from P to P_c elements, and so do the fused clip/SGD pass, the weight traffic of those convolutions
        (9x) and the broadcast / gather passes.  Returns ``None`` when nothing can be elided, else
        ``{offsets, compact, numel, index_map}`` with ``index_map[j]`` = position of slot element j in the global arena
        (-1 = alignment padding)."""
        shapes = {}
        hooks = []
        for name, mod in template.net.named_modules():
            if isinstance(mod, torch.nn.Conv2d):
                hooks.append(mod.register_forward_pre_hook(
                    lambda m, inp, n=name: shapes.__setitem__(n, tuple(inp[0].shape[-2:]))))
        try:
            with torch.no_grad():
                template.net(example_input[:2].float())
        finally:
            for h in hooks:
                h.remove()
        mods = dict(template.net.named_modules())
        names = [n for n, _ in template.named_parameters()]
        offsets, compact, parts, cur = {}, [], [], 0
        for n, o, k, sh in zip(names, layout.offsets, layout.sizes, layout.shapes):
            idx = None
            mod_name = n[len("net."):-len(".weight")] if n.startswith("net.") and n.endswith(".weight") else None
            if mod_name in shapes and len(sh) == 4:
                conv = mods[mod_name]
                KH, KW = int(sh[2]), int(sh[3])
                taps = live_taps(shapes[mod_name][0], shapes[mod_name][1], KH, KW, conv.stride[0], conv.padding[0])
                if len(taps) < KH * KW and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]:
                    base = torch.arange(int(sh[0]) * int(sh[1]), dtype=torch.int64).view(-1, 1) * (KH * KW)
                    tap = torch.tensor([kh * KW + kw for kh, kw in taps], dtype=torch.int64).view(1, -1)
                    idx = (o + base + tap).reshape(-1)
                    compact.append(n)
            if idx is None:
                idx = torch.arange(o, o + k, dtype=torch.int64)
            pad = (-cur) % 32                                   # 128-byte aligned tensors, like the global arena
            if pad:
                parts.append(torch.full((pad,), -1, dtype=torch.int64))
                cur += pad
            offsets[n] = cur
            parts.append(idx)
            cur += idx.numel()
        if not compact:
            return None
        pad = (-cur) % 32
        if pad:
            parts.append(torch.full((pad,), -1, dtype=torch.int64))
            cur += pad
        return {"offsets": offsets, "compact": compact, "numel": cur,
                "index_map": torch.cat(parts).to(torch.int32)}

    # -- layer helpers (x is [S, B, C, H, W]) ---------------------------------------------------------------
    def _conv(self, x, name, conv):
        pname = "net." + name + ".weight"
        return SlotConv2d.apply(x, self.dummy, self.W, self.G, self.off[pname], conv.out_channels,
                                conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0],
                                pname in self.compact)

    def _gn(self, x, name, gn, residual=None, relu=False):
        S, B = x.shape[0], x.shape[1]
        x4 = x.reshape((S * B,) + tuple(x.shape[2:]))
        r4 = residual.reshape(x4.shape) if residual is not None else None
        y = SlotGroupNorm.apply(x4, r4, self.dummy, self.W, self.G, self.off["net." + name + ".weight"],
                                self.off["net." + name + ".bias"], gn.num_groups, gn.eps, relu, True, S)
        return y.view(x.shape)

    def _block(self, x, prefix, blk):
        if blk.downsample is None:
            residual = x
        else:
            r = self._conv(x, prefix + ".downsample.0", getattr(blk.downsample, "0"))
            residual = self._gn(r, prefix + ".downsample.1", getattr(blk.downsample, "1"))
        if isinstance(blk, BasicBlock):
            out = self._gn(self._conv(x, prefix + ".conv1", blk.conv1), prefix + ".bn1", blk.bn1, relu=True)
            return self._gn(self._conv(out, prefix + ".conv2", blk.conv2), prefix + ".bn2", blk.bn2, residual=residual,
                            relu=True)
        out = self._gn(self._conv(x, prefix + ".conv1", blk.conv1), prefix + ".bn1", blk.bn1, relu=True)
        out = self._gn(self._conv(out, prefix + ".conv2", blk.conv2), prefix + ".bn2", blk.bn2, relu=True)
        return self._gn(self._conv(out, prefix + ".conv3", blk.conv3), prefix + ".bn3", blk.bn3, residual=residual,
                        relu=True)

    def logits(self, x):
        """x: [S, B, 3, H, W] fp32 → [S, B, num_classes]."""
        net = self.net
        S, B = x.shape[0], x.shape[1]
        x = self._gn(self._conv(x, "conv1", net.conv1), "bn1", net.bn1, relu=True)
        pool = misc_ops.max_pool2d if _SLOT_MAXPOOL else F.max_pool2d     # hand-written K8 kernel (1-byte argmax)
        x4 = pool(x.reshape((S * B,) + tuple(x.shape[2:])), 3, 2, 1)
        x = x4.view((S, B) + tuple(x4.shape[1:]))
        for li, layer in enumerate((net.layer1, net.layer2, net.layer3, net.layer4), start=1):
            for bi, blk in enumerate(layer):
                x = self._block(x, "layer{}.{}".format(li, bi), blk)
        feat = x.reshape(S, B, -1)
        fc = net.fc
        return SlotLinear.apply(feat, self.dummy, self.W, self.G, self.off["net.fc.weight"], self.off["net.fc.bias"],
                                fc.out_features, fc.in_features)

    def losses(self, x, y):
        """Per-slot mean cross-entropy [S] (the sum of which is back-propagated)."""
        logits = self.logits(x)
        S, B = y.shape
        ce = misc_ops.softmax_cross_entropy(logits.reshape(S * B, -1), y.reshape(-1))      # one fused fwd+bwd kernel
        return ce.view(S, B).mean(dim=1)
