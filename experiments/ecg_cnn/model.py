"""ecg_cnn task model."""
from msrflute_b200.models.ecg import SuperNet, Net, ConvNormPool, RNN, Swish  # noqa: F401
