"""Minimal stand-in for `fvcore.common.config.CfgNode` (yacs) so the reference's config-driven constructors and
`utils/experiment_manager.py` work without fvcore/yacs/iopath (not installed here).

Behaviour kept (utils/experiment_manager.py:11-35, fvcore load_yaml_with_base):
  * attribute access on nested dicts, every node `new_allowed`;
  * `_BASE_: "other.yaml"` inheritance, relative to the including file, base first then overlay;
  * `merge_from_list(["KEY.SUB", "value", ...])` with `ast.literal_eval` of string values;
  * string leaves that parse as Python literals are converted on merge (yacs `_decode_cfg_value`), which is what
    turns PyYAML's `LR: 1e-4` string into a float.
"""
from __future__ import annotations

import ast
import copy
import os
import sys
import types
from typing import Any

import yaml

BASE_KEY = "_BASE_"


def _decode(v: Any) -> Any:
    if isinstance(v, dict) and not isinstance(v, CfgNode):
        return CfgNode(v)
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


class CfgNode(dict):
    NEW_ALLOWED = "__new_allowed__"
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__[CfgNode.NEW_ALLOWED] = True
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = type(self)(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else _decode(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get(CfgNode.IMMUTABLE):
            raise AttributeError(f"Attempted to set {name} on an immutable CfgNode")
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        self.__dict__[CfgNode.IMMUTABLE] = True

    def defrost(self):
        self.__dict__[CfgNode.IMMUTABLE] = False

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    # -- yaml -------------------------------------------------------------------------------------------
    @classmethod
    def load_yaml_with_base(cls, filename: str, allow_unsafe: bool = False) -> dict:
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) if not allow_unsafe else yaml.unsafe_load(f)
        cfg = cfg or {}

        def merge_a_into_b(a: dict, b: dict) -> None:
            for k, v in a.items():
                if isinstance(v, dict) and isinstance(b.get(k), dict):
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if BASE_KEY in cfg:
            base = cfg.pop(BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not base.startswith("/"):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = cls.load_yaml_with_base(base, allow_unsafe=allow_unsafe)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename: str, allow_unsafe: bool = True) -> None:
        self.merge_from_other_cfg(type(self)(self.load_yaml_with_base(cfg_filename, allow_unsafe=allow_unsafe)))

    def merge_from_other_cfg(self, other: "CfgNode") -> None:
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_other_cfg(v if isinstance(v, CfgNode) else type(self)(v))
            else:
                self[k] = copy.deepcopy(v) if isinstance(v, CfgNode) else _decode(v)

    def merge_from_list(self, cfg_list) -> None:
        if len(cfg_list) % 2 != 0:
            raise AssertionError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split(".")
            for sub in keys[:-1]:
                if sub not in node:
                    node[sub] = type(self)()
                node = node[sub]
            node[keys[-1]] = _decode(v)


def install_fvcore_stub() -> None:
    """Registers `fvcore.common.config.CfgNode` in sys.modules unless the real fvcore is importable."""
    try:
        import fvcore.common.config  # noqa: F401
        return
    except Exception:  # noqa: BLE001
        pass
    fv = types.ModuleType("fvcore")
    common = types.ModuleType("fvcore.common")
    config = types.ModuleType("fvcore.common.config")
    config.CfgNode = CfgNode
    fv.common = common
    common.config = config
    sys.modules["fvcore"] = fv
    sys.modules["fvcore.common"] = common
    sys.modules["fvcore.common.config"] = config


def new_config() -> CfgNode:
    """Same pre-created groups as utils/experiment_manager.py:38-56."""
    C = CfgNode()
    C.CONFIG_DIR = "config/"
    for k in ("PATHS", "TRAINER", "MODEL", "DATALOADER", "AUGMENTATIONS", "CONSISTENCY_TRAINER", "DATASETS"):
        C[k] = CfgNode()
    return C.clone()


def load_cfg(yaml_path: str, opts=()) -> CfgNode:
    """Load a reference YAML (with `_BASE_` chain) plus `KEY VALUE` overrides, as `setup_cfg` does
    (utils/experiment_manager.py:59-69) minus the path assertions."""
    cfg = new_config()
    cfg.merge_from_file(str(yaml_path))
    cfg.merge_from_list(list(opts))
    cfg.NAME = os.path.splitext(os.path.basename(str(yaml_path)))[0]
    return cfg


def synthetic_cfg(model_type: str, in_channels: int = 6, topology=(64, 128, 256, 512), s1_bands=(0, 1),
                  s2_bands=(2, 1, 0, 3), **extra) -> CfgNode:
    """Config carrying exactly the keys the network constructors read (utils/networks.py:64-66,91,96,105)."""
    cfg = new_config()
    cfg.SEED = 7
    cfg.MODEL.TYPE = model_type
    cfg.MODEL.IN_CHANNELS = in_channels
    cfg.MODEL.OUT_CHANNELS = 1
    cfg.MODEL.TOPOLOGY = list(topology)
    cfg.MODEL.LOSS_TYPE = "PowerJaccardLoss"
    cfg.DATALOADER.S1_BANDS = list(s1_bands)
    cfg.DATALOADER.S2_BANDS = list(s2_bands)
    cfg.CONSISTENCY_TRAINER.LOSS_FACTOR = 0.5
    cfg.CONSISTENCY_TRAINER.LOSS_TYPE = "PowerJaccardLoss"
    cfg.TRAINER.LR = 1e-4            # configs/base.yaml:8 (read by load_checkpoint, utils/networks.py:47)
    cfg.TRAINER.BATCH_SIZE = 8       # configs/base.yaml:9
    for k, v in extra.items():
        cfg[k] = v
    return cfg
