"""Logistic regression for LR-MNIST (ref. ``experiments/cv_lr_mnist/model.py:12-36``): ``Linear(784,10)``
followed by a sigmoid, then cross-entropy on the sigmoid outputs (a FedML quirk that is kept)."""
import torch

from .common import ClassifierModel


class LogisticRegression(torch.nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim = input_dim
        self.linear = torch.nn.Linear(input_dim, output_dim)

    def forward(self, x):
        return torch.sigmoid(self.linear(x.reshape(-1, self.input_dim).float()))


class LR(ClassifierModel):
    def __init__(self, model_config):
        super().__init__(model_config)
        self.net = LogisticRegression(model_config["input_dim"], model_config["output_dim"])
