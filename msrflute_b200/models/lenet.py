"""LeNet-style CIFAR-10 classifier with an extra ``f1_score`` metric (ref. ``experiments/classif_cnn/model.py:11-62``).
The micro-F1 is computed on the device (for single-label multi-class micro-F1 == accuracy) instead of a
``sklearn`` call on host copies every batch."""
import torch
from torch import nn
from torch.nn import functional as F

from .common import ClassifierModel


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 6, 5)
        self.pool = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(6, 16, 5)
        self.fc1 = nn.Linear(16 * 5 * 5, 120)
        self.fc2 = nn.Linear(120, 84)
        self.fc3 = nn.Linear(84, 10)

    def forward(self, x):
        x = self.pool(F.relu(self.conv1(x.float())))
        x = self.pool(F.relu(self.conv2(x)))
        x = torch.flatten(x, 1)
        return self.fc3(F.relu(self.fc2(F.relu(self.fc1(x)))))


def micro_f1(pred, labels, num_classes):
    """micro-averaged F1 = Σ TP / (Σ TP + ½(Σ FP + Σ FN)), on the device."""
    tp = (pred == labels).sum().float()
    fp_fn = 2.0 * (pred != labels).sum().float()        # every error is one FP and one FN
    return tp / (tp + 0.5 * fp_fn).clamp(min=1.0)


class CNN(ClassifierModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.net = Net()

    def inference(self, input):
        x, y = self._xy(input)
        logits = self.forward(x)
        pred = logits.argmax(1)
        return {"output": logits, "acc": (pred == y).float().mean().item(), "batch_size": x.shape[0],
                "f1_score": {"value": micro_f1(pred, y, 10).item(), "higher_is_better": True}}

    def loss_and_metrics(self, input):
        x, y = self._xy(input)
        logits = self.forward(x)
        pred = logits.argmax(1)
        return F.cross_entropy(logits, y), {"output": None, "acc": (pred == y).float().mean(), "batch_size": x.shape[0],
                                            "f1_score": {"value": micro_f1(pred, y, 10), "higher_is_better": True}}
