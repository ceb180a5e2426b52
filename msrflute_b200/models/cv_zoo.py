"""ResNet / ResNeXt / WideResNet / VGG classifiers for the personalization scenario
(ref. ``experiments/cv/model.py:160-489`` and ``model_vgg.py`` — torchvision-style networks re-declared as
``BaseModel`` subclasses).  Here the torchvision constructors are used directly and wrapped once.

Reference behaviour kept: inputs arrive as ``batch['x']`` in (N, H, W, C)-like layout and are transposed with
``transpose(1, 3)`` before the stem (ref :244); ``inference`` returns ``output = {'probabilities' (log-softmax),
'predictions', 'labels'}`` so ``convex_inference`` can mix local and global models (ref :288-303).
"""
import numpy as np
import torch
import torch.nn.functional as F
import torchvision

from ..core.model import BaseModel


class _TVClassifier(BaseModel):
    arch = "resnet18"

    def __init__(self, model_config, **kwargs):
        super().__init__()
        num_classes = int(model_config.get("num_classes", 10)) if hasattr(model_config, "get") else 10
        self.net = getattr(torchvision.models, self.arch)(weights=None, num_classes=num_classes)

    def _dev(self):
        return next(self.parameters()).device

    def forward(self, inputs):
        x = inputs["x"] if isinstance(inputs, dict) else inputs
        x = x.to(self._dev(), non_blocking=True).float()
        return self.net(torch.transpose(x, 1, 3))

    def loss(self, inputs):
        y = inputs["y"].to(self._dev(), non_blocking=True).long()
        self.train()
        return F.cross_entropy(self.forward(inputs), y)

    def inference(self, inputs):
        y = inputs["y"].to(self._dev(), non_blocking=True).long()
        self.eval()
        logp = F.log_softmax(self.forward(inputs), dim=1)
        acc = (logp.argmax(1) == y).float().mean().item()
        out = {"probabilities": logp.detach().cpu().numpy(), "predictions": np.arange(0, y.shape[0]),
               "labels": y.cpu().numpy()}
        return {"output": out, "acc": acc, "batch_size": y.shape[0]}

    def get_logit(self, x=None, evalis=True, logmax=False):
        data, target = x
        fn = F.log_softmax if logmax else F.softmax
        if evalis:
            self.eval()
            with torch.no_grad():
                logits = fn(self.forward(data), dim=1)
        else:
            self.train()
            logits = fn(self.forward(data), dim=1)
        return logits.cpu(), target.cpu(), 1


def _make(name):
    return type(name, (_TVClassifier,), {"arch": name, "__doc__": "torchvision ``{}`` as a FLUTE model".format(name)})


resnet18, resnet34, resnet50 = _make("resnet18"), _make("resnet34"), _make("resnet50")
resnet101, resnet152 = _make("resnet101"), _make("resnet152")
resnext50_32x4d, resnext101_32x8d = _make("resnext50_32x4d"), _make("resnext101_32x8d")
wide_resnet50_2, wide_resnet101_2 = _make("wide_resnet50_2"), _make("wide_resnet101_2")
vgg11, vgg11_bn, vgg13, vgg13_bn = _make("vgg11"), _make("vgg11_bn"), _make("vgg13"), _make("vgg13_bn")
vgg16, vgg16_bn, vgg19, vgg19_bn = _make("vgg16"), _make("vgg16_bn"), _make("vgg19"), _make("vgg19_bn")
