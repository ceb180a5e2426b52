"""ResNet-18 FedCIFAR-100 task model (BASELINE config #3, the headline benchmark)."""
from msrflute_b200.models.resnet_gn import (RESNET, ResNet, BasicBlock, Bottleneck, GroupNorm2d,  # noqa: F401
                                            resnet18, resnet34, resnet50, resnet101, resnet152)
