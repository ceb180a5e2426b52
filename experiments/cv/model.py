"""cv (personalization) task models: ``model_type`` ∈ resnet18…152, resnext*, wide_resnet*."""
from msrflute_b200.models.cv_zoo import (resnet18, resnet34, resnet50, resnet101, resnet152,  # noqa: F401
                                         resnext50_32x4d, resnext101_32x8d, wide_resnet50_2, wide_resnet101_2)
