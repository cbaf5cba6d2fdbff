"""YAML -> attribute dict config with `_BASE_CONFIG_` includes and `--set` overrides.

Same surface as the reference's pcdet/config.py:7-85 (cfg, cfg_from_yaml_file, cfg_from_list,
log_config_to_file, merge_new_config) without the easydict dependency.
"""
from __future__ import annotations

import ast
from pathlib import Path

import yaml


class EasyDict(dict):
    """dict with attribute access; nested dicts (also inside lists) are converted on assignment."""

    def __init__(self, d=None, **kw):
        super().__init__()
        self.update(d or {}, **kw)

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return EasyDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(EasyDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def update(self, d=None, **kw):
        for k, v in dict(d or {}, **kw).items():
            self[k] = v


def log_config_to_file(cfg, pre="cfg", logger=None):
    for key, val in cfg.items():
        if isinstance(val, EasyDict):
            logger.info("\n%s.%s = edict()" % (pre, key))
            log_config_to_file(val, pre=f"{pre}.{key}", logger=logger)
        else:
            logger.info("%s.%s: %s" % (pre, key, val))


def cfg_from_list(cfg_list, config):
    """`--set A.B.C value ...`: the key must exist; the value is coerced to the existing type
    (`k:v,k:v` for dict-valued keys, comma lists for list-valued keys)."""
    assert len(cfg_list) % 2 == 0
    for dotted, raw in zip(cfg_list[0::2], cfg_list[1::2]):
        *parents, leaf = dotted.split(".")
        d = config
        for p in parents:
            assert p in d, "NotFoundKey: %s" % p
            d = d[p]
        assert leaf in d, "NotFoundKey: %s" % leaf
        try:
            value = ast.literal_eval(raw)
        except (ValueError, SyntaxError):
            value = raw
        old = d[leaf]
        if type(value) is not type(old) and isinstance(old, EasyDict):
            for item in str(raw).split(","):
                k, v = item.split(":")
                old[k] = type(old[k])(v)
        elif type(value) is not type(old) and isinstance(old, list):
            d[leaf] = [type(old[0])(x) for x in str(raw).split(",")]
        else:
            assert type(value) is type(old), f"type {type(value)} does not match original type {type(old)}"
            d[leaf] = value


def merge_new_config(config, new_config):
    base = new_config.get("_BASE_CONFIG_")
    if base is not None:
        with open(base) as f:                      # relative to the CWD (run from tools/, like the reference)
            config.update(EasyDict(yaml.safe_load(f)))
    for key, val in new_config.items():
        if not isinstance(val, dict):
            config[key] = val
            continue
        if key not in config:
            config[key] = EasyDict()
        merge_new_config(config[key], val)
    return config


def cfg_from_yaml_file(cfg_file, config):
    with open(cfg_file) as f:
        merge_new_config(config=config, new_config=yaml.safe_load(f))
    return config


cfg = EasyDict()
cfg.ROOT_DIR = (Path(__file__).resolve().parent / "../").resolve()
cfg.LOCAL_RANK = 0
