"""fednewsrec task model."""
from msrflute_b200.models.newsrec import FEDNEWS, FedNewsRec  # noqa: F401
