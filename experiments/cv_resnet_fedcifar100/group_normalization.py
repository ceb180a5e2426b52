from msrflute_b200.models.resnet_gn import GroupNorm2d  # noqa: F401
