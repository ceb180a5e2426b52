"""Offline stand-in for ``cerberus``: the dependency-free validator of this repo implements the rule subset the
reference schema uses (required / type / allowed / default / nullable / allow_unknown / schema / keysrules)."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(__file__), "..", "..", "..", "msrflute_b200", "core", "validator.py")
_spec = importlib.util.spec_from_file_location("_flute_validator", os.path.abspath(_p))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
Validator = _m.Validator
SchemaError = _m.SchemaError
