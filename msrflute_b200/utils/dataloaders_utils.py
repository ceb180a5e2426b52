"""Task plug-in resolution for datasets/dataloaders.

Same convention as the reference (``utils/dataloaders_utils.py:9-23,85-98``):
``experiments/<task>/dataloaders/dataloader.py`` must define ``DataLoader`` and
``experiments/<task>/dataloaders/dataset.py`` must define ``Dataset``.  Paths
are searched relative to the working directory first and then relative to the
repository that contains this package, so jobs can be launched from anywhere.
Loaded modules are cached (the reference re-executes the file on every call,
i.e. once per simulated client).
"""
import importlib.util
import os
import sys

from .utils import print_rank

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_MODULE_CACHE = {}


def resolve_path(rel_path):
    for base in (os.getcwd(), _REPO_ROOT):
        p = os.path.join(base, rel_path)
        if os.path.exists(p):
            return p
    raise FileNotFoundError("{} not found under {} or {}".format(rel_path, os.getcwd(), _REPO_ROOT))


def load_source(name, path):
    path = os.path.abspath(path)
    if path in _MODULE_CACHE:
        return _MODULE_CACHE[path]
    root = os.path.dirname(os.path.dirname(os.path.dirname(path)))  # .../experiments/<task>/x.py -> repo root guess
    for cand in (_REPO_ROOT, root):
        if cand not in sys.path:
            sys.path.insert(0, cand)
    spec = importlib.util.spec_from_file_location("_flute_plugin_" + name + "_" + str(len(_MODULE_CACHE)), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _MODULE_CACHE[path] = mod
    return mod


def get_exp_dataloader(task):
    path = resolve_path(os.path.join("experiments", task, "dataloaders", "dataloader.py"))
    return load_source("DataLoader", path).DataLoader


def get_exp_dataset(task):
    path = resolve_path(os.path.join("experiments", task, "dataloaders", "dataset.py"))
    return load_source("Dataset", path).Dataset


def make_train_dataloader(data_config, data_path, clientx, task=None, vec_size=300, data_strct=None,
                          replay_server=False):
    """Client (``clientx`` = int) or server-replay (``clientx`` = None) train loader."""
    mode = "train"
    tokenizer_type = data_config.get("tokenizer_type", "not_applicable")
    if clientx is None:
        if not data_config.get("train_data_server", None):
            print_rank("No server training set is defined")
            return None
        my_data = os.path.join(data_path, data_config["train_data_server"])
        mode, clientx = "val", 0
    elif tokenizer_type != "not_applicable" and "train_data" in data_config:
        my_data = data_config["train_data"][clientx]
    else:
        my_data = data_config.get("list_of_train_data", None)
    loader_cls = get_exp_dataloader(task)
    return loader_cls(data=data_strct if data_strct is not None else my_data, user_idx=clientx, mode=mode,
                      args=data_config)


def _eval_loader(mode, key, data_config, data_path, task, data_strct):
    f = data_config.get(key, None)
    path = os.path.join(data_path, f) if f is not None and data_path is not None else None
    return get_exp_dataloader(task)(data=data_strct if data_strct is not None else path, user_idx=0, mode=mode,
                                    args=data_config)


def make_val_dataloader(data_config, data_path, task=None, data_strct=None, train_mode=False):
    return _eval_loader("val", "val_data", data_config, data_path, task, data_strct)


def make_test_dataloader(data_config, data_path, task=None, data_strct=None):
    return _eval_loader("test", "test_data", data_config, data_path, task, data_strct)


def get_data_config(config, mode):
    if mode in ("val", "test"):
        dc = config["server_config"]["data_config"][mode]
    else:
        dc = config["client_config"]["data_config"]["train"]
    semi = config["client_config"].get("semisupervision", None)
    return dc if semi is None else {**dc, **semi}


def get_dataset(data_path, config, task, mode, test_only=False, user_idx=-1, data_strct=None):
    dc = get_data_config(config, mode)
    key = {"val": "val_data", "test": "test_data"}.get(mode, "list_of_train_data")
    f = dc.get(key, None)
    pointer = os.path.join(data_path, f) if (f is not None and data_path is not None) else f
    return get_exp_dataset(task)(pointer if data_strct is None else data_strct, test_only=test_only,
                                 user_idx=user_idx, args=dc)
