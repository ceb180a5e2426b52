"""classif_cnn task model (LeNet on CIFAR-10 + custom f1 metric)."""
from msrflute_b200.models.lenet import CNN, Net  # noqa: F401
