"""Config plumbing of the reference's plugin boundary, on plain dicts + PyYAML.

`instantiate_from_config` mirrors /root/reference/utils/util.py:45-63: import `config['target']`
(a dotted path) and call it with `**config['params']` (+ merged extras).  OmegaConf is not a
dependency here; yaml is read with `yaml.safe_load` (note: it parses `lr: 5e-5` as a string --
only training would care)."""
from __future__ import annotations

import importlib

import yaml


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config, merge=None):
    if "target" not in config:
        if config == "__is_first_stage__" or config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = dict(config.get("params", dict()) or {})
    if merge is not None:
        params.update(dict(merge))  # key sets never overlap in the reference's callers (mage_model.py:475-477)
    return get_obj_from_str(config["target"])(**params)


def load_yaml(path: str) -> dict:
    with open(path, "r") as fp:
        return yaml.safe_load(fp)
