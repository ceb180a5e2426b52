"""Configuration objects.

Parity target: ``core/config.py`` of the reference — a tree of objects that
behave like nested dicts (``cfg['a']['b']``, ``.get``, ``in``, ``pop``) and
also support dotted ``lookup('a.b.c')`` (``config.py:39-79``), built from a
YAML dict after schema validation + default filling (``config.py:762-796``)
and post-processed by ``validate()`` which joins data paths
(``config.py:736-760``).

Design here: one generic ``ConfigNode`` (a ``dict`` subclass with attribute
access) instead of ~15 hand-written dataclasses; the named classes the
reference exposes are kept as thin subclasses so ``isinstance`` checks and
``XConfig.from_dict`` calls keep working.  Because a ``ConfigNode`` *is* a
dict it deep-copies, pickles (gloo object collectives) and YAML-dumps without
any special casing.
"""
from __future__ import annotations

import copy
import os
from importlib.machinery import SourceFileLoader
from typing import Any, Mapping

from .schema import SCHEMA
from .validator import Validator


class ConfigNode(dict):
    """dict with attribute access, ``None``-aware ``get`` and dotted lookup."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    # -- construction -----------------------------------------------------
    @classmethod
    def _wrap(cls, v):
        if isinstance(v, ConfigNode):
            return v
        if isinstance(v, Mapping):
            return ConfigNode(v)
        return v

    @classmethod
    def from_dict(cls, config: Mapping):
        return cls(config) if config is not None else None

    def to_dict(self) -> dict:
        def conv(v):
            if isinstance(v, Mapping):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [conv(x) for x in v]
            return v
        return conv(self)

    # -- mapping behaviour ------------------------------------------------
    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __contains__(self, k):
        # reference semantics (config.py:69-70): a key holding None is "absent"
        return dict.__contains__(self, k) and dict.__getitem__(self, k) is not None

    def get(self, k, default=None):
        v = dict.get(self, k, default)
        return default if v is None else v

    def pop(self, k, default=None):
        v = self.get(k, default)
        if dict.__contains__(self, k):
            dict.__delitem__(self, k)
        return v

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def lookup(self, path: str, default=None):
        node: Any = self
        for tok in path.split("."):
            if isinstance(node, Mapping) and dict.__contains__(node, tok) and node[tok] is not None:
                node = node[tok]
            else:
                return default
        return node

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def __reduce__(self):
        return (type(self), (dict(self),))


# Named aliases kept for API parity with the reference's dataclass tree.
class Config(ConfigNode):
    pass


class ModelConfig(Config):
    @staticmethod
    def from_dict(config):
        """Honour the reference's plug-in hook (``config.py:100-116``): if a
        ``config.py`` sits next to the model file and defines
        ``<model_type>Config``, build that class; otherwise keep a plain node."""
        folder = str(config.get("model_folder", ""))
        cfg_path = os.path.join(os.path.dirname(os.path.join(".", folder)), "config.py")
        if os.path.exists(cfg_path):
            mod = SourceFileLoader("_task_config", cfg_path).load_module()
            klass = getattr(mod, str(config["model_type"]) + "Config", None)
            if klass is not None:
                try:
                    return klass.from_dict(config) if hasattr(klass, "from_dict") else klass(**config)
                except TypeError:
                    pass
        return ModelConfig(config)


class BERTModelConfig(Config): pass
class BERTTrainingConfig(Config): pass
class BERTConfig(Config): pass
class PrivacyConfig(Config): pass
class PrivacyMetricsConfig(Config): pass
class OptimizerConfig(Config): pass
class AnnealingConfig(Config): pass
class DatasetConfig(Config): pass
class DataConfig(Config): pass
class ServerReplayConfig(Config): pass
class RLConfig(Config): pass
class ServerConfig(Config): pass
class ClientConfig(Config): pass


def _retag(node, klass):
    return klass(node) if isinstance(node, Mapping) and not isinstance(node, klass) else node


class FLUTEConfig(Config):
    """Root configuration of a training job."""

    SECTIONS = ("model_config", "dp_config", "privacy_metrics_config", "strategy",
                "server_config", "client_config")

    @staticmethod
    def from_dict(config: Mapping, lenient_numbers: bool = True) -> "FLUTEConfig":
        v = Validator(SCHEMA, allow_unknown=True, lenient_numbers=lenient_numbers)
        if not v.validate(config):
            raise ValueError("Missing {} argumment in config file ".format(v.errors))
        cfg = FLUTEConfig(v.normalized(config))
        cfg.setdefault("strategy", "DGA")
        cfg["model_config"] = ModelConfig.from_dict(cfg["model_config"])
        cfg["dp_config"] = _retag(cfg.get("dp_config"), PrivacyConfig)
        cfg["privacy_metrics_config"] = _retag(cfg.get("privacy_metrics_config"), PrivacyMetricsConfig)
        srv = cfg["server_config"] = _retag(cfg["server_config"], ServerConfig)
        cli = cfg["client_config"] = _retag(cfg["client_config"], ClientConfig)
        for sec in (srv, cli):
            if "optimizer_config" in sec:
                sec["optimizer_config"] = _retag(sec["optimizer_config"], OptimizerConfig)
            if "annealing_config" in sec:
                sec["annealing_config"] = _retag(sec["annealing_config"], AnnealingConfig)
            sec["data_config"] = _retag(sec["data_config"], DataConfig)
            for mode in list(sec["data_config"].keys()):
                sec["data_config"][mode] = _retag(sec["data_config"][mode], DatasetConfig)
        if "server_replay_config" in srv:
            srv["server_replay_config"] = _retag(srv["server_replay_config"], ServerReplayConfig)
        if "RL" in srv:
            srv["RL"] = _retag(srv["RL"], RLConfig)
        return cfg

    def validate(self) -> "FLUTEConfig":
        """Join relative paths with ``data_path``/``output_path`` (ref. config.py:736-760)."""
        data_path = self.get("data_path", "") or ""
        srv = self["server_config"]
        if srv.get("wantRL", False) and "RL" in srv and "RL_path" in srv["RL"]:
            rl = srv["RL"]
            base = self.get("output_path", "") if rl.get("RL_path_global", True) else \
                os.path.join(self.get("output_path", ""), self.get("experiment_name", ""))
            rl["RL_path"] = os.path.join(base, rl["RL_path"])
        mc = self["model_config"]
        if "pretrained_model_path" in mc:
            mc["pretrained_model_path"] = os.path.join(data_path, mc["pretrained_model_path"])
        for section in ("server_config", "client_config"):
            dc = self[section]["data_config"]
            for mode in ("test", "val", "train"):
                if mode not in dc:
                    continue
                if "vocab_dict" in dc[mode]:
                    dc[mode]["vocab_dict"] = os.path.join(data_path, dc[mode]["vocab_dict"])
                if "BERT" in mc:
                    bm = mc["BERT"]["model"]
                    want_section = "client_config" if mode == "train" else "server_config"
                    if section == want_section:
                        dc[mode]["model_name_or_path"] = bm["model_name"]
                        dc[mode]["process_line_by_line"] = bm["process_line_by_line"]
        return self

    @staticmethod
    def from_yaml(path: str) -> "FLUTEConfig":
        import yaml
        with open(path) as f:
            return FLUTEConfig.from_dict(yaml.safe_load(f))
