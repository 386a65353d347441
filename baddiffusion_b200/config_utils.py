"""Minimal config surface compatible with the vendored diffusers' ConfigMixin JSON files
(D/configuration_utils.py: `config.json` / `scheduler_config.json`, sorted keys, `_class_name`,
`_diffusers_version`).  Only what the BadDiffusion hot path needs: save, load, attribute access, mutation
(`model.py:640` assigns `noise_sched.config.clip_sample`)."""
from __future__ import annotations

import inspect
import json
import os

DIFFUSERS_VERSION = "0.16.0.dev0"  # D/__init__.py:1 -- what the reference writes into its JSON files


class Config(dict):
    """dict with attribute access (read and write)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def capture_init_args(obj, init, args, kwargs, ignore=()):
    """Equivalent of @register_to_config (D/configuration_utils.py:549-591): all ctor args incl. defaults."""
    sig = inspect.signature(init)
    params = [p for n, p in sig.parameters.items() if n != "self"]
    cfg = Config()
    for p in params:
        if p.default is not inspect.Parameter.empty:
            cfg[p.name] = p.default
    for p, a in zip(params, args):
        cfg[p.name] = a
    cfg.update({k: v for k, v in kwargs.items() if k in sig.parameters})
    for k in ignore:
        cfg.pop(k, None)
    for k, v in list(cfg.items()):
        if isinstance(v, list):
            cfg[k] = tuple(v)
    obj.config = cfg
    return cfg


def to_json_string(cfg: dict, class_name: str) -> str:
    d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    d["_class_name"] = class_name
    d["_diffusers_version"] = DIFFUSERS_VERSION
    return json.dumps(d, indent=2, sort_keys=True) + "\n"


def save_config(cfg: dict, class_name: str, directory: str, filename: str):
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, filename), "w", encoding="utf-8") as f:
        f.write(to_json_string(cfg, class_name))


def load_config(directory: str, filename: str) -> dict:
    path = os.path.join(directory, filename)
    if not os.path.isfile(path):
        raise EnvironmentError(f"Error no file named {filename} found in directory {directory}.")
    with open(path, "r", encoding="utf-8") as f:
        d = json.load(f)
    return d


def filter_init_kwargs(cls, d: dict) -> dict:
    """from_config semantics (D/configuration_utils.py:421-470): keep only the keys the ctor accepts."""
    sig = inspect.signature(cls.__init__)
    return {k: v for k, v in d.items() if k in sig.parameters and not k.startswith("_")}
