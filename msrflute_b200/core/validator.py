"""A small, dependency-free schema validator/normaliser.

The reference validates its YAML with the third-party ``cerberus`` package
(``core/config.py:762-773``: ``Validator(schema).validate`` then
``.normalized`` to fill defaults).  ``cerberus`` is not a dependency here; this
module implements the subset of its rule language that the FL config schema
needs: ``type``, ``required``, ``nullable``, ``allowed``, ``default``,
``schema`` (nested mapping), ``allow_unknown`` and ``keysrules.forbidden``.

The class is deliberately API-compatible with ``cerberus.Validator`` for the
calls the reference makes (``validate(doc, schema)``, ``errors``,
``normalized(doc)``) so it can also stand in for it when the unmodified
reference is run as the benchmark baseline (see ``baseline/shims``).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Mapping, Optional

_TYPE_CHECKS = {
    "string": lambda v: isinstance(v, str),
    "boolean": lambda v: isinstance(v, bool),
    # cerberus: bool is not an integer/float, ints are not floats ...
    "integer": lambda v: isinstance(v, int) and not isinstance(v, bool),
    # ... but YAML users routinely write ``lr: 1`` — cerberus rejects it, we
    # follow cerberus ("float" accepts float only) unless lenient=True.
    "float": lambda v: isinstance(v, float),
    "number": lambda v: isinstance(v, (int, float)) and not isinstance(v, bool),
    "dict": lambda v: isinstance(v, Mapping),
    "list": lambda v: isinstance(v, (list, tuple)),
}


class SchemaError(ValueError):
    """Raised when the schema itself is malformed."""


class Validator:
    """Validate and normalise nested mappings against a rule schema."""

    def __init__(self, schema: Optional[Dict[str, Any]] = None, allow_unknown: bool = False,
                 lenient_numbers: bool = False):
        self.schema = schema
        self.allow_unknown = allow_unknown
        self.lenient_numbers = lenient_numbers
        self.errors: Dict[str, Any] = {}

    # ------------------------------------------------------------------ API
    def validate(self, document: Mapping, schema: Optional[Dict[str, Any]] = None) -> bool:
        schema = schema if schema is not None else self.schema
        if schema is None:
            raise SchemaError("no schema given")
        self.errors = {}
        self._check_mapping(document, schema, self.allow_unknown, self.errors)
        return not self.errors

    def normalized(self, document: Mapping, schema: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Return a deep copy of ``document`` with schema defaults filled in."""
        schema = schema if schema is not None else self.schema
        out = copy.deepcopy(dict(document))
        self._fill_defaults(out, schema)
        return out

    # ------------------------------------------------------------ internals
    def _type_ok(self, rule_type, value) -> bool:
        types = rule_type if isinstance(rule_type, (list, tuple)) else [rule_type]
        for t in types:
            if t not in _TYPE_CHECKS:
                raise SchemaError(f"unknown type rule {t!r}")
            if _TYPE_CHECKS[t](value):
                return True
            if t == "float" and self.lenient_numbers and _TYPE_CHECKS["integer"](value):
                return True
        return False

    def _check_mapping(self, doc, schema, allow_unknown, errors):
        if not isinstance(doc, Mapping):
            errors["__self__"] = ["must be of dict type"]
            return
        for key, rules in schema.items():
            if key not in doc:
                if rules.get("required", False) and "default" not in rules:
                    errors[key] = ["required field"]
                continue
            self._check_value(key, doc[key], rules, errors)
        if not allow_unknown:
            for key in doc:
                if key not in schema:
                    errors[key] = ["unknown field"]

    def _check_value(self, key, value, rules, errors):
        errs = []
        if value is None:
            if not rules.get("nullable", False):
                errs.append("null value not allowed")
            if errs:
                errors[key] = errs
            return
        if "type" in rules and not self._type_ok(rules["type"], value):
            errs.append(f"must be of {rules['type']} type")
            errors[key] = errs
            return
        if "allowed" in rules and value not in rules["allowed"]:
            errs.append(f"unallowed value {value}")
        if isinstance(value, Mapping):
            forbidden = (rules.get("keysrules") or {}).get("forbidden", [])
            bad = [k for k in value if k in forbidden]
            if bad:
                errs.append({k: ["forbidden key"] for k in bad})
            if "schema" in rules:
                sub = {}
                self._check_mapping(value, rules["schema"], rules.get("allow_unknown", self.allow_unknown), sub)
                if sub:
                    errs.append(sub)
        if errs:
            errors[key] = errs

    def _fill_defaults(self, doc, schema):
        for key, rules in schema.items():
            if key not in doc:
                if "default" in rules:
                    doc[key] = copy.deepcopy(rules["default"])
                else:
                    continue
            if isinstance(doc[key], Mapping) and "schema" in rules:
                doc[key] = dict(doc[key])
                self._fill_defaults(doc[key], rules["schema"])
