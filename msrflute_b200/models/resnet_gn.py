"""ResNet family with GroupNorm (or BatchNorm) for FedCIFAR-100 (ref. ``experiments/cv_resnet_fedcifar100/model.py``
— a port of FedML's ``resnet_gn.py`` — and ``group_normalization.py``).

Topology is torchvision-style: 7×7/2 stem, 3×3/2 max-pool, four stages of BasicBlock/Bottleneck,
``AvgPool2d(1)`` (identity; with 32×32 inputs the last stage is already 1×1), FC.  Conv weights ~ N(0, √(2/n)),
norm weights 1 except the last norm of every residual branch which starts at 0 (ref :115-134).

``group_norm`` = channels per group (FedML uses 2); 0 selects ``BatchNorm2d``.  As *shipped* the reference's
``RESNET`` wrapper calls ``resnet18()`` with defaults, i.e. BatchNorm and a 1000-way head (ref :253) —
``RESNET`` below reproduces that when ``model_config`` gives no overrides, and the benchmark config sets
``group_norm: 2`` as BASELINE.json names the GroupNorm variant.

``GroupNorm2d`` keeps the reference's semantics: statistics over (channels-in-group × H × W) per sample, and an
affine pair **per group**, not per channel (ref ``group_normalization.py:59-84``).  On CUDA it dispatches to the
fused hand-written kernel (``ops.norm_ops.group_norm``: normalise + affine + optional residual-add + ReLU in one
pass, fwd and bwd).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops import norm_ops
from .common import ClassifierModel


class GroupNorm2d(nn.Module):
    def __init__(self, num_features, channels_per_group, eps=1e-5, affine=True):
        super().__init__()
        assert num_features % channels_per_group == 0
        self.num_features, self.channels_per_group = num_features, channels_per_group
        self.num_groups = num_features // channels_per_group
        self.eps = eps
        if affine:
            self.weight = nn.Parameter(torch.ones(self.num_groups))
            self.bias = nn.Parameter(torch.zeros(self.num_groups))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x, residual=None, relu=False):
        return norm_ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, residual=residual,
                                   relu=relu, per_group_affine=True)

    def extra_repr(self):
        return "{}, groups={}, eps={}".format(self.num_features, self.num_groups, self.eps)


class _BN(nn.BatchNorm2d):
    """BatchNorm2d — what the reference's RESNET instantiates as shipped (model.py:116).  Training-mode forward /
    backward with the residual add and ReLU fused run on ``csrc/nn_kernels.cu`` (one CTA per channel); evaluation
    (running statistics) is a plain affine map."""

    def forward(self, x, residual=None, relu=False):
        if self.training and x.is_cuda and x.dtype == torch.float32 and self.affine and self.track_running_stats:
            from ..ops import nn_ops
            if self.momentum is not None:
                if self.num_batches_tracked is not None:
                    self.num_batches_tracked.add_(1)
                return nn_ops.batch_norm_train(x, self.weight, self.bias, self.running_mean, self.running_var,
                                               self.momentum, self.eps, residual=residual, relu=relu)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y


def norm2d(planes, channels_per_group=0):
    return GroupNorm2d(planes, channels_per_group) if channels_per_group > 0 else _BN(planes)


def conv3x3(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1, bias=False)


class _Down(nn.Module):
    def __init__(self, inplanes, outplanes, stride, gn):
        super().__init__()
        self.add_module("0", nn.Conv2d(inplanes, outplanes, kernel_size=1, stride=stride, bias=False))
        self.add_module("1", norm2d(outplanes, gn))

    def forward(self, x):
        return getattr(self, "1")(getattr(self, "0")(x))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, group_norm=0):
        super().__init__()
        self.conv1, self.bn1 = conv3x3(inplanes, planes, stride), norm2d(planes, group_norm)
        self.conv2, self.bn2 = conv3x3(planes, planes), norm2d(planes, group_norm)
        self.downsample = downsample

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x), relu=True)
        return self.bn2(self.conv2(out), residual=residual, relu=True)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, group_norm=0):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), norm2d(planes, group_norm)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False), norm2d(planes, group_norm)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), norm2d(planes * 4, group_norm)
        self.downsample = downsample

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x), relu=True)
        out = self.bn2(self.conv2(out), relu=True)
        return self.bn3(self.conv3(out), residual=residual, relu=True)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, group_norm=0):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm2d(64, group_norm)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], 1, group_norm)
        self.layer2 = self._make_layer(block, 128, layers[1], 2, group_norm)
        self.layer3 = self._make_layer(block, 256, layers[2], 2, group_norm)
        self.layer4 = self._make_layer(block, 512, layers[3], 2, group_norm)
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
        for m in self.modules():
            if isinstance(m, Bottleneck):
                m.bn3.weight.data.fill_(0)
            elif isinstance(m, BasicBlock):
                m.bn2.weight.data.fill_(0)

    def _make_layer(self, block, planes, blocks, stride, gn):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = _Down(self.inplanes, planes * block.expansion, stride, gn)
        layers = [block(self.inplanes, planes, stride, down, gn)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, group_norm=gn) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.bn1(self.conv1(x), relu=True))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(x, 1))


_DEPTHS = {"resnet18": (BasicBlock, [2, 2, 2, 2]), "resnet34": (BasicBlock, [3, 4, 6, 3]),
           "resnet50": (Bottleneck, [3, 4, 6, 3]), "resnet101": (Bottleneck, [3, 4, 23, 3]),
           "resnet152": (Bottleneck, [3, 8, 36, 3])}


def build_resnet(name="resnet18", **kwargs):
    block, layers = _DEPTHS[name]
    return ResNet(block, layers, **kwargs)


def resnet18(**kw): return build_resnet("resnet18", **kw)
def resnet34(**kw): return build_resnet("resnet34", **kw)
def resnet50(**kw): return build_resnet("resnet50", **kw)
def resnet101(**kw): return build_resnet("resnet101", **kw)
def resnet152(**kw): return build_resnet("resnet152", **kw)


class RESNET(ClassifierModel):
    """``model_config`` keys (all optional): ``arch`` (resnet18), ``group_norm`` (0 ⇒ BatchNorm — the reference's
    shipped default; 2 = FedML GroupNorm), ``num_classes`` (1000, the reference's shipped default), ``channels_last``."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self.net = build_resnet(model_config.get("arch", "resnet18"),
                                num_classes=int(model_config.get("num_classes", 1000)),
                                group_norm=int(model_config.get("group_norm", 0)))

    def forward(self, x):
        if x.dtype == torch.uint8:
            x = x.float()
        return super().forward(x)
