"""CNN-FEMNIST task model (BASELINE config #2)."""
from msrflute_b200.models.cnn_femnist import CNN, CNN_DropOut  # noqa: F401
