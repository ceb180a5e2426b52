"""semisupervision (FedLabels) task model: CIFAR-100 ResNet (ref. ``experiments/semisupervision/model.py``: ``Res`` wraps
a CIFAR-style ResNet-50 with BatchNorm; here ``arch`` is configurable — resnet50 by default, resnet18 for smoke tests).
``loss`` / ``inference`` take ``{'x','y'}`` batches of already-normalised CHW images; ``forward(images)`` (called by the
FedLabels trainer directly) returns logits."""
import torch
import torch.nn.functional as F
import torchvision

from msrflute_b200.core.model import BaseModel


class Res(BaseModel):
    def __init__(self, model_config):
        super().__init__()
        arch = model_config.get("arch", "resnet50")
        net = getattr(torchvision.models, arch)(weights=None, num_classes=int(model_config.get("num_classes", 100)))
        net.conv1 = torch.nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)    # CIFAR stem
        net.maxpool = torch.nn.Identity()
        self.net = net

    def forward(self, x):
        return self.net(x.to(next(self.parameters()).device).float())

    def loss(self, input):
        y = input["y"].to(next(self.parameters()).device).long()
        return F.cross_entropy(self.forward(input["x"]), y)

    def inference(self, input):
        y = input["y"].to(next(self.parameters()).device).long()
        out = self.forward(input["x"])
        return {"output": out, "acc": (out.argmax(1) == y).float().mean().item(), "batch_size": y.shape[0]}
