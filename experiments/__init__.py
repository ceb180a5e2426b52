"""Task plug-ins.  ``make_model`` re-exported for reference compatibility (``from experiments import make_model``)."""
from msrflute_b200.models import make_model  # noqa: F401
