"""``server_config.type: personalization`` server (the reference keeps it in the task folder)."""
from msrflute_b200.core.server import PersonalizationServer  # noqa: F401
