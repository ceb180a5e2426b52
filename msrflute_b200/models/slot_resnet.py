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
from ..ops.slot_ops import SlotConv2d, SlotGroupNorm, SlotLinear
from .resnet_gn import BasicBlock, Bottleneck, GroupNorm2d, RESNET


class SlotBatchedResNet:
    @staticmethod
    def supports(model) -> bool:
        if not isinstance(model, RESNET) or next(model.parameters()).dtype != torch.float32:
            return False
        if getattr(model, "compute_dtype", "fp32") != "fp32":
            return False
        return not any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in model.modules())

    def __init__(self, template: RESNET, layout, W, G):
        self.net = template.net
        self.W, self.G, self.S = W, G, W.shape[0]
        names = [n for n, _ in template.named_parameters()]
        self.off = {n: o for n, o in zip(names, layout.offsets)}
        self.dummy = torch.zeros(1, device=W.device, requires_grad=True)

    # -- layer helpers (x is [S, B, C, H, W]) ---------------------------------------------------------------
    def _conv(self, x, name, conv):
        return SlotConv2d.apply(x, self.dummy, self.W, self.G, self.off["net." + name + ".weight"], conv.out_channels,
                                conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0])

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
        x4 = F.max_pool2d(x.reshape((S * B,) + tuple(x.shape[2:])), 3, 2, 1)
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
