"""cv (personalization) task models: VGG family."""
from msrflute_b200.models.cv_zoo import vgg11, vgg11_bn, vgg13, vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn  # noqa: F401
