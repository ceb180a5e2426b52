"""FEMNIST CNN (ref. ``experiments/cv_cnn_femnist/model.py:53-80``; "Adaptive Federated Optimization" model):
conv3×3(1→32) → ReLU → conv3×3(32→64) → ReLU → maxpool2 → dropout .25 → FC 9216→128 → ReLU → dropout .5 →
FC 128→{10|62}.  1,206,590 parameters with 62 classes."""
import torch
from torch import nn

from ..ops.nn_ops import Dropout
from .common import ClassifierModel


class CNN_DropOut(nn.Module):
    def __init__(self, only_digits=True):
        super().__init__()
        self.conv2d_1 = nn.Conv2d(1, 32, kernel_size=3)
        self.conv2d_2 = nn.Conv2d(32, 64, kernel_size=3)
        self.max_pooling = nn.MaxPool2d(2, stride=2)
        self.dropout_1 = Dropout(0.25)          # Philox mask recomputed in the backward (csrc/nn_kernels.cu)
        self.linear_1 = nn.Linear(9216, 128)
        self.dropout_2 = Dropout(0.5)
        self.linear_2 = nn.Linear(128, 10 if only_digits else 62)

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        x = torch.relu(self.conv2d_1(x.float()))
        x = torch.relu(self.conv2d_2(x))
        x = self.dropout_1(self.max_pooling(x))
        x = self.dropout_2(torch.relu(self.linear_1(torch.flatten(x, 1))))
        return self.linear_2(x)


class CNN(ClassifierModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.net = CNN_DropOut(bool(model_config.get("only_digits", False)))
